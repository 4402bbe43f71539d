#!/bin/bash
# One GPU call of a development session: full -m gpu suite, A/B bench lines, launch list, ncu captures.
# usage (from the repo root, under gpurun): bash tools/gpu_session.sh <tag> [stages...]   stages: test bench ab launches ncu
TAG=${1:-dev}; shift
STAGES=${@:-test bench ab launches ncu}
mkdir -p gpurun_out
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has test; then
  (time timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 240 -p no:cacheprovider 2>&1 | tail -150) > gpurun_out/${TAG}_pytest.log 2>&1
  tail -4 gpurun_out/${TAG}_pytest.log
fi
if has bench; then
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  tail -c 1500 gpurun_out/${TAG}_bench.json
fi
if has ab; then
  NKSR_FILL=rows NKSR_SPMV=rows NKSR_ROWS=location timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mesh > gpurun_out/${TAG}_bench_legacy.json 2> gpurun_out/${TAG}_bench_legacy.err
  NKSR_COMPACT_ROWS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mesh > gpurun_out/${TAG}_bench_compact.json 2> gpurun_out/${TAG}_bench_compact.err
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py cfg4_outdoor_10M mesh > gpurun_out/${TAG}_launches.log 2>&1
fi
if has ncu; then
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_gram_fill_group|k_spmv_stream|k_build_rows_voxel|k_knn_normals" -c 30 -o gpurun_out/${TAG}_hot python tools/profile_run.py cfg4_outdoor_10M > gpurun_out/${TAG}_ncu.log 2>&1
fi
ls -la gpurun_out | tail -20
