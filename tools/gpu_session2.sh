#!/bin/bash
# Multi-GPU session (N = $1): partitioned parity at a production-like shape, CG with / without PDL + trace, full bench.
N=${1:-2}; TAG=${2:-m}; shift 2
STEPS=${@:-test dist cgab bench}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PLANES=$((64 * N))
for s in $STEPS; do
  case $s in
    test)
      python -m pytest tests/test_krylov_gpu.py tests/test_blas_cg_gpu.py tests/test_dist_gpu.py -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${TAG}_pytest.log ;;
    dist)
      timeout 1200 $TR --master-port 29541 tools/dist_check.py --production > $O/${TAG}_dist_check.json 2> $O/${TAG}_dist_check.err; echo "dist_check rc=$?"
      python -c "
import json
d=json.loads([l for l in open('$O/${TAG}_dist_check.json') if l.startswith('{')][-1]); print({k:v for k,v in d.items() if k!='cases'}); print('cases ok:', all(all(c[k]==1 for k in ('spmv_exact','cg_same_iters','cg_hist','cg_x')) for c in d['cases']), len(d['cases']))" ;;
    cgab)
      for mode in 1 0; do
        B200SP_CG_PDL=$mode B200SP_CG_TRACE=$O/${TAG}_trace_pdl$mode timeout 600 $TR --master-port 29542 bench.py --gpus $N --only-cg --cg-grid 512x512x$PLANES --steps 10 --warmup 3 > $O/${TAG}_cg_pdl$mode.json 2> $O/${TAG}_cg_pdl$mode.err; echo "cg pdl=$mode rc=$?"
        python -c "
import json
d=json.loads(open('$O/${TAG}_cg_pdl$mode.json').read().strip().splitlines()[-1]); print('pdl=$mode', d['cg'], d['parity'])"
        python tools/cg_trace.py $O/${TAG}_trace_pdl$mode > $O/${TAG}_trace_pdl$mode.summary.json; cat $O/${TAG}_trace_pdl$mode.summary.json | head -24
      done ;;
    bench1)
      timeout 900 python bench.py > $O/${TAG}_bench1.json 2> $O/${TAG}_bench1.err; echo "bench1 rc=$?"; tail -c 2500 $O/${TAG}_bench1.json ;;
    bench)
      timeout 900 $TR --master-port 29543 bench.py --gpus $N > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1500 $O/${TAG}_bench.json ;;
  esac
done
