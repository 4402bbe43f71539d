#!/usr/bin/env bash
# GPU-side half of the experiments/ protocol (see experiments/README.md).  Before the gpurun call, locally:
#     git apply experiments/r2_all.patch && make -C nksr_b200/csrc        (the built .so travels to the box)
# then:   gpurun --timeout 600 -- 'bash tools/run_experiments.sh'
# Writes gpurun_out/exp_tests.log and one JSON line per (switch, value) under gpurun_out/exp_<switch>_<v>.json.
set -u
mkdir -p gpurun_out
NKSR_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_exp_*.py -m gpu -q --tb=short > gpurun_out/exp_tests.log 2>&1
tail -5 gpurun_out/exp_tests.log
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_depths.py tests/test_golden.py -m gpu -q --tb=short \
    > gpurun_out/exp_parity.log 2>&1
tail -3 gpurun_out/exp_parity.log
run() {   # name, env assignment
  env "$2" NKSR_BENCH_WATCHDOG=80 timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
      > "gpurun_out/exp_$1.json" 2> /dev/null
  python - "$1" <<'EOF'
import json, sys
try:
    d = json.load(open(f"gpurun_out/exp_{sys.argv[1]}.json"))
    st = d["solve"]["stages_ms_profiled_run"]
    print(sys.argv[1], round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["ms_per_step"], 1), "spmv_frac",
          round(d["roofline"]["frac"], 3), {k: round(v, 1) for k, v in st.items()})
except Exception as e:          # a failed variant must not hide the others
    print(sys.argv[1], "FAILED", e)
EOF
}
run base        NKSR_NONE=0
run flush       NKSR_GROUPED_FLUSH=1
run writeout    NKSR_MASKED_WRITEOUT=1
run spmv        NKSR_SPMV_PIPELINE=1
run e2e         NKSR_E2E_OVERLAP=1
