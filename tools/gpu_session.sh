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
  NKSR_SPMV=stream timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mesh > gpurun_out/${TAG}_bench_stream.json 2> gpurun_out/${TAG}_bench_stream.err
  fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_run.py cfg4_outdoor_10M 10000000 mesh > gpurun_out/${TAG}_launches.log 2>&1
fi
if has ncu; then
  # gpurun_out/ must stay under 64 MiB: export the pages that are read afterwards and drop the report
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_gram_fill" -c 1 -o /tmp/${TAG}_fill python tools/profile_run.py cfg4_outdoor_10M > gpurun_out/${TAG}_ncu.log 2>&1
  ncu -i /tmp/${TAG}_fill.ncu-rep --page raw --csv > gpurun_out/${TAG}_fill_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_fill.ncu-rep --page source --csv --print-source sass > /tmp/${TAG}_fill_source.csv 2>/dev/null
  python tools/hot_lines.py /tmp/${TAG}_fill_source.csv > gpurun_out/${TAG}_fill_hot_sass.txt 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:"k_spmv_stream|k_knn_normals|k_gram_blocks|k_place_rank|k_place_prefix" -c 24 -o /tmp/${TAG}_other python tools/profile_run.py cfg4_outdoor_10M >> gpurun_out/${TAG}_ncu.log 2>&1
  ncu -i /tmp/${TAG}_other.ncu-rep --page raw --csv > gpurun_out/${TAG}_other_raw.csv 2>/dev/null
fi
du -sh gpurun_out
ls -la gpurun_out | tail -20
if has spmvncu; then
  timeout 600 ncu --set full --clock-control none -k regex:"k_spmv_stream|k_spmv" -c 6 -o /tmp/${TAG}_spmv python tools/profile_run.py cfg4_outdoor_10M > gpurun_out/${TAG}_spmv_ncu.log 2>&1
  ncu -i /tmp/${TAG}_spmv.ncu-rep --page raw --csv > gpurun_out/${TAG}_spmv_raw.csv 2>/dev/null
fi
if has global; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_dcg|k_gather|k_scatter|k_reduce|k_spmv" -c 200 --csv --log-file gpurun_out/${TAG}_global_launches.csv python tools/profile_global.py > gpurun_out/${TAG}_global.log 2>&1
  timeout 300 python tools/profile_global.py > gpurun_out/${TAG}_global_plain.log 2>&1
fi
if has final; then
  (time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${TAG}_pytest.log 2>&1
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log | tail -2; tail -c 600 gpurun_out/${TAG}_bench_reference.json
fi
if has split1; then
  NKSR_BLOCK_SPLIT=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mesh > gpurun_out/${TAG}_bench_split1.json 2> gpurun_out/${TAG}_bench_split1.err
fi
if has w1; then
  # r2w: the new kernels' own tests, then A/B bench lines of the interleaved row layout, the overlapped count and the U-Net
  (time timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_parity.py -q --tb=short --timeout 240 -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/${TAG}_pytest_new.log 2>&1
  tail -3 gpurun_out/${TAG}_pytest_new.log
  B="--steps 3 --warmup 3 --no-cpu-baseline --no-mesh"
  timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_base.json 2> gpurun_out/${TAG}_bench_base.err
  NKSR_ROW_LAYOUT=interleaved timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_ilv.json 2> gpurun_out/${TAG}_bench_ilv.err
  NKSR_ROW_LAYOUT=interleaved NKSR_OVERLAP=1 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_ilv_overlap.json 2> gpurun_out/${TAG}_bench_ilv_overlap.err
  NKSR_OVERLAP=1 timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_overlap.json 2> gpurun_out/${TAG}_bench_overlap.err
  timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mesh --backbone unet-tf32 > gpurun_out/${TAG}_bench_unet_tf32.json 2> gpurun_out/${TAG}_bench_unet_tf32.err
  timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mesh --backbone unet > gpurun_out/${TAG}_bench_unet.json 2> gpurun_out/${TAG}_bench_unet.err
  for f in base ilv ilv_overlap overlap unet_tf32 unet; do echo $f; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_$f.json") if l.startswith("{")][-1]
    print(d["ms_per_step"], d["solve"]["stages_ms_timed_steps"], d.get("hbm_peak_allocated_gb"))
except Exception as e:
    print("no line:", e); print(open("gpurun_out/${TAG}_bench_$f.err").read()[-1500:])
PY
  done
fi
if has w2; then
  (time timeout 600 python -m pytest tests/test_gpu_parity.py -q --tb=short --timeout 240 -p no:cacheprovider -k "interleaved" 2>&1 | tail -30) > gpurun_out/${TAG}_pytest_ilv.log 2>&1
  tail -3 gpurun_out/${TAG}_pytest_ilv.log
  B="--steps 3 --warmup 3 --no-cpu-baseline --no-mesh"
  NKSR_ROW_LAYOUT=interleaved timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_ilv.json 2> gpurun_out/${TAG}_bench_ilv.err
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_ilv.json") if l.startswith("{")][-1]
print(d["ms_per_step"], d["solve"]["stages_ms_timed_steps"], d.get("hbm_peak_allocated_gb"))
PY
fi
if has w3; then
  (time timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py tests/test_gpu_api_surface.py -q --tb=short --timeout 300 -p no:cacheprovider -k "chunk or reconstruct or distributed or api or surface" 2>&1 | tail -60) > gpurun_out/${TAG}_pytest_chunk.log 2>&1
  tail -5 gpurun_out/${TAG}_pytest_chunk.log
fi
if has w4; then
  (time timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_pipeline.py -q --tb=short --timeout 300 -p no:cacheprovider -k "gather or unet or chunk" 2>&1 | tail -60) > gpurun_out/${TAG}_pytest_w4.log 2>&1
  tail -5 gpurun_out/${TAG}_pytest_w4.log
  timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mesh --backbone unet-tf32 > gpurun_out/${TAG}_bench_unet_tf32.json 2> gpurun_out/${TAG}_bench_unet_tf32.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_unet_tf32.json") if l.startswith("{")][-1]
    print(d["ms_per_step"], d["solve"]["stages_ms_timed_steps"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/${TAG}_bench_unet_tf32.err").read()[-1500:])
PY
fi
if has y1; then
  # r2y: tcgen05 gather-GEMM diagnostics first (own process: a trap there must not take the suite down), then the
  # whole -m gpu suite without the tcgen05 tests, smoke, the tcgen05 tests, the headline bench, the U-Net bench on tcgen05
  timeout 240 python tools/tc_diag.py ${TAG} > gpurun_out/${TAG}_tc_diag.log 2>&1; echo "tc_diag rc=$?"; tail -5 gpurun_out/${TAG}_tc_diag.log | cut -c1-400
  (time timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 240 -p no:cacheprovider -k "not tcgen05" 2>&1 | tail -40) > gpurun_out/${TAG}_pytest.log 2>&1
  tail -4 gpurun_out/${TAG}_pytest.log
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
  (timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q --tb=short --timeout 120 -p no:cacheprovider -k "tcgen05" 2>&1 | tail -40) > gpurun_out/${TAG}_pytest_tcgen05.log 2>&1
  tail -3 gpurun_out/${TAG}_pytest_tcgen05.log
  timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  tail -c 700 gpurun_out/${TAG}_bench.json
  timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mesh --backbone unet-tc > gpurun_out/${TAG}_bench_unet_tc.json 2> gpurun_out/${TAG}_bench_unet_tc.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_unet_tc.json") if l.startswith("{")][-1]
    print("unet-tc", d["ms_per_step"], d["solve"]["stages_ms_timed_steps"])
except Exception as e:
    print("no unet-tc line", e); print(open("gpurun_out/${TAG}_bench_unet_tc.err").read()[-800:])
PY
fi
if has y2; then
  timeout 200 python tools/chunk_diag.py > gpurun_out/${TAG}_chunk_diag.log 2>&1; tail -2 gpurun_out/${TAG}_chunk_diag.log | cut -c1-900
  (timeout 200 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short --timeout 150 -p no:cacheprovider -k "chunked" 2>&1 | tail -25) > gpurun_out/${TAG}_pytest_chunk.log 2>&1
  tail -3 gpurun_out/${TAG}_pytest_chunk.log
  timeout 240 python tools/tc_diag.py ${TAG} > gpurun_out/${TAG}_tc_diag_tn64.log 2>&1; grep -E "time_|unet_" gpurun_out/${TAG}_tc_diag_tn64.log | cut -c1-420
  NKSR_TC_TN=128 timeout 240 python tools/tc_diag.py ${TAG} > gpurun_out/${TAG}_tc_diag_tn128.log 2>&1; grep -E "time_|unet_|random27" gpurun_out/${TAG}_tc_diag_tn128.log | cut -c1-420
  NKSR_TC_UNET=0 timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_gather_gemm_t" --launch-skip 79 -c 9 -o /tmp/${TAG}_tc python tools/tc_diag.py ${TAG}ncu > gpurun_out/${TAG}_tc_ncu.log 2>&1
  ncu -i /tmp/${TAG}_tc.ncu-rep --page raw --csv > gpurun_out/${TAG}_tc_ncu_raw.csv 2>/dev/null
  ls -la gpurun_out | tail -12
fi
if has y3; then
  (time timeout 420 python -m pytest tests -m gpu -q --tb=short --timeout 200 -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/${TAG}_pytest_final.log 2>&1
  tail -4 gpurun_out/${TAG}_pytest_final.log
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke_final.log 2>&1; tail -1 gpurun_out/${TAG}_smoke_final.log
  timeout 150 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mesh --backbone unet-tc > gpurun_out/${TAG}_bench_unet_tc2.json 2> gpurun_out/${TAG}_bench_unet_tc2.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_unet_tc2.json") if l.startswith("{")][-1]
    print("unet-tc", d["ms_per_step"], d["solve"]["stages_ms_timed_steps"])
except Exception as e:
    print("no unet-tc line", e); print(open("gpurun_out/${TAG}_bench_unet_tc2.err").read()[-800:])
PY
fi
if has y4; then
  NKSR_TC_QUICK=1 timeout 70 python tools/tc_diag.py ${TAG}q > gpurun_out/${TAG}_tc_quick.log 2>&1; grep -E "time_|unet_" gpurun_out/${TAG}_tc_quick.log | cut -c1-400
  NKSR_TC_QUICK=1 NKSR_TC_UNET=0 timeout 60 ncu --set full --import-source on --clock-control none -k regex:"k_gather_gemm_tc" --launch-skip 2 -c 2 -o /tmp/${TAG}_tc128 python tools/tc_diag.py ${TAG}ncu2 > gpurun_out/${TAG}_tc128_ncu.log 2>&1
  ncu -i /tmp/${TAG}_tc128.ncu-rep --page raw --csv > gpurun_out/${TAG}_tc128_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_tc128.ncu-rep --page source --csv --print-source sass > /tmp/${TAG}_tc128_source.csv 2>/dev/null
  python tools/hot_lines.py /tmp/${TAG}_tc128_source.csv > gpurun_out/${TAG}_tc128_hot_sass.txt 2>&1
fi
