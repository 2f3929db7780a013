#!/bin/bash
# One-GPU session: full GPU test suite, bench line, COO sweep, launch list + --set full captures of the final kernels.
# usage (on the GPU box, from the repo root): bash tools/gpu_session.sh <tag> [steps...]   steps: test bench probe ncu
TAG=${1:-s}; shift
STEPS=${@:-test bench probe ncu}
O=gpurun_out
mkdir -p $O
for s in $STEPS; do
  case $s in
    test)
      python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log ;;
    bench)
      python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1800 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err ;;
    cgtest)
      timeout 900 python -m pytest tests/test_blas_cg_gpu.py tests/test_krylov_gpu.py -m gpu -x -q > $O/${TAG}_cgtest.log 2>&1; echo "cgtest rc=$?"; tail -6 $O/${TAG}_cgtest.log ;;
    cgab)
      for f in 0 1; do
        B200SP_CG_FUSE=$f timeout 600 python bench.py --only-cg --steps 10 --warmup 3 > $O/${TAG}_cg_fuse$f.json 2> $O/${TAG}_cg_fuse$f.err; echo "cg fuse=$f rc=$?"
        python -c "
import json
d=json.loads(open('$O/${TAG}_cg_fuse$f.json').read().strip().splitlines()[-1]); print('fuse=$f', d['cg'], d['parity'])"
      done ;;
    runab)
      for run in 1 4 16 64; do
        B200SP_DIA_RUN=$run timeout 600 python bench.py --only-cg --steps 10 --warmup 3 > $O/${TAG}_cg_run$run.json 2> $O/${TAG}_cg_run$run.err; echo "cg run=$run rc=$?"
        python -c "
import json
d=json.loads(open('$O/${TAG}_cg_run$run.json').read().strip().splitlines()[-1]); print('fused cg run=$run', d['cg']['ms_per_iter'], d['parity']['cg_ok'])"
      done
      for run in 1 4 16; do
        B200SP_DIA_RUN=$run timeout 600 python bench.py --quick --steps 100 > $O/${TAG}_dia_run$run.json 2> $O/${TAG}_dia_run$run.err; echo "dia run=$run rc=$?"
        python -c "
import json
d=json.loads(open('$O/${TAG}_dia_run$run.json').read().strip().splitlines()[-1]); print('plain dia run=$run', d['ms_per_step'], d['roofline']['frac'])"
      done ;;
    ncucg)
      python tools/prof_kernels.py cg > $O/${TAG}_cgplain.log 2>&1 && \
      ncu --set full --clock-control none --import-source on -k regex:"dia_bulk|cg_update|cg_direction" -c 8 \
          -f -o $O/${TAG}_cg_full python tools/prof_kernels.py cg > $O/${TAG}_ncu_cg.log 2>&1; echo "ncu cg rc=$?"
      B200SP_CG_FUSE=0 python tools/prof_kernels.py cg > $O/${TAG}_cgplain3.log 2>&1 && \
      B200SP_CG_FUSE=0 ncu --set full --clock-control none --import-source on -k regex:"dia_bulk|cg_update|cg_direction" -c 10 \
          -f -o $O/${TAG}_cg3_full python tools/prof_kernels.py cg > $O/${TAG}_ncu_cg3.log 2>&1; echo "ncu cg3 rc=$?" ;;
    shapeab)
      for shape in 128,2,2,4 256,1,2,4 256,1,3,4 256,1,2,6 128,2,3,4 256,2,2,2 128,2,2,3; do
        B200SP_DIA_RUN=1 B200SP_DIA_FUSED_SHAPE=$shape timeout 600 python bench.py --only-cg --steps 10 --warmup 3 > $O/${TAG}_cg_shape.json 2> $O/${TAG}_cg_shape.err; echo "cg shape=$shape rc=$?"
        python -c "
import json
d=json.loads(open('$O/${TAG}_cg_shape.json').read().strip().splitlines()[-1]); print('fused cg shape=$shape', d['cg']['ms_per_iter'], d['parity']['cg_ok'])"
      done ;;
    smallcg)
      python - > $O/${TAG}_smallcg.log 2>&1 <<'PY'
import os, time, torch
import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import gallery
h = cusp.default_handle()
dev = torch.device("cuda", 0)
for grid in ((512, 512), (256, 256), (128, 128)):
    M = gallery.poisson("csr", 5, grid, dtype=torch.float64)
    n = M.num_rows
    b = torch.ones(n, dtype=torch.float64, device=dev)
    for env in ({"B200SP_CG_PERSISTENT": "0", "B200SP_CG_GRAPH": "0"}, {"B200SP_CG_PERSISTENT": "0"}, {}):
        os.environ.pop("B200SP_CG_PERSISTENT", None); os.environ.pop("B200SP_CG_GRAPH", None)
        os.environ.update(env)
        for rep in range(2):
            x = torch.zeros(n, dtype=torch.float64, device=dev)
            torch.cuda.synchronize(); t = time.perf_counter()
            r, _ = h.cg(M.descriptor(), x, b, iteration_limit=1024, relative_tolerance=0.0, check_interval=128, want_residuals=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(grid, env, "us/iter", round(dt / max(1, int(r.iteration_count)) * 1e6, 2), "iters", int(r.iteration_count), "res", r.residual_norm)
PY
      echo "smallcg rc=$?"; cat $O/${TAG}_smallcg.log ;;
    nculist)
      python bench.py --quick --steps 20 --warmup 3 > $O/${TAG}_quick.json 2> $O/${TAG}_quick.err && \
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench_quick.csv \
          python bench.py --quick --steps 20 --warmup 3 > $O/${TAG}_ncu_quick.log 2>&1; echo "ncu bench list rc=$?"
      python tools/prof_kernels.py cg > $O/${TAG}_cgplain.log 2>&1 && \
      ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_launches_cg.csv \
          python tools/prof_kernels.py cg > $O/${TAG}_ncu_cg_list.log 2>&1; echo "ncu cg list rc=$?" ;;
    refcoo)
      REF_COO=full timeout 1500 python tools/ref_gpu_kernels.py 256 8 > $O/${TAG}_ref_kernels.json 2> $O/${TAG}_ref_kernels.err; echo "refcoo rc=$?"
      python -c "
import json
d=json.loads(open('$O/${TAG}_ref_kernels.json').read().strip().splitlines()[-1]); c=d['coo']
print({k: c[k] for k in c if k not in ('reference_top','invalid')}); print([ (r['cfg'], round(r['ms'],4)) for r in c['reference_top'][:6]]); print(c['invalid'][:3])"
      tail -3 $O/${TAG}_ref_kernels.err ;;
    e2eab)
      for ch in 16 32 64; do for aff in 0 1; do
        B200SP_HOST_CHUNKS=$ch B200SP_BENCH_AFFINITY=$aff timeout 300 python bench.py --quick --steps 50 > $O/${TAG}_e2e.json 2> $O/${TAG}_e2e.err
        python -c "
import json
d=json.loads(open('$O/${TAG}_e2e.json').read().strip().splitlines()[-1]); print('chunks=$ch affinity=$aff', d['e2e']['ms_per_step'], d['e2e'].get('host_affinity'), d['parity'])"
      done; done ;;
    hybprobe)
      timeout 600 python tools/hyb_probe.py > $O/${TAG}_hyb_probe.json 2> $O/${TAG}_hyb_probe.err; echo "hyb probe rc=$?"; cat $O/${TAG}_hyb_probe.json; tail -3 $O/${TAG}_hyb_probe.err ;;
    widen)
      timeout 900 python tools/widen_time.py > $O/${TAG}_widen.json 2> $O/${TAG}_widen.err; echo "widen rc=$?"; tail -c 3000 $O/${TAG}_widen.json; tail -3 $O/${TAG}_widen.err ;;
    benchref)
      python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "benchref rc=$?" ;;
    probe)
      python tools/coo_probe.py 24 > $O/${TAG}_coo_probe.log 2>&1; echo "probe rc=$?"; grep -A9 "^## scale" $O/${TAG}_coo_probe.log ;;
    ncu)
      python tools/prof_kernels.py dia,ell,csr,coop,coo,plan 2 > $O/${TAG}_plain.log 2>&1 && \
      ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv \
          python tools/prof_kernels.py dia,ell,csr,coop,coo,plan 2 > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
      python tools/prof_kernels.py dia,ell,csr,coop,coo,plan 2 > $O/${TAG}_plain2.log 2>&1 && \
      ncu --set full --clock-control none --import-source on -k regex:"dia_bulk|ell_bulk|csr_ring|coo_ring|coo_warp" -c 12 \
          -f -o $O/${TAG}_full python tools/prof_kernels.py dia,ell,csr,coop,coo,plan 2 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?" ;;
  esac
done
