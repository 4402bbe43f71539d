"""A/B of the Gram-assembly variants on one workload (dev tool)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench, nksr_b200
wl = sys.argv[1] if len(sys.argv) > 1 else "dev_outdoor_1M"
pts = int(sys.argv[2]) if len(sys.argv) > 2 else None
dev = torch.device("cuda:0")
xyz, sensor = bench.make_cloud(wl, 4, 0, points=pts)
xyz, sensor = xyz.to(dev), sensor.to(dev)
prep = nksr_b200.get_estimate_normal_preprocess_fn(64, 85.0)
os.environ["NKSR_STAGE_TIMES"] = "1"
orig_init = nksr_b200.fields.KernelField.__init__
for variant, compact in [("row", True), ("row", False), ("group", True), ("group", False)]:
    os.environ["NKSR_FILL_VARIANT"] = variant
    def init(self, *a, _c=compact, **k):
        orig_init(self, *a, **k)
        self.solver_config["compact_rows"] = _c
    nksr_b200.fields.KernelField.__init__ = init
    rec = nksr_b200.Reconstructor(dev)
    for rep in range(2):
        f = rec.reconstruct(xyz, sensor=sensor, voxel_size=bench.WORKLOADS[wl]["voxel_size"], preprocess_fn=prep, **bench.SOLVER)
        del f
    st = rec.last_stats["stages_ms"]
    print(variant, "compact" if compact else "3row", {k: round(v, 1) for k, v in st.items()}, "peakGB", round(torch.cuda.max_memory_allocated() / 1e9, 1), flush=True)
