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
