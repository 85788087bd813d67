"""Device timeline of inference chunks (config 5 path: NeRFSystem.forward(train=False), 4096-ray chunks)."""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
system = bench.make_system("bf16", dev)
system.set_progress(0.75)
system.hparams["val.chunk_size"] = 4096
n = 4096 * 6
b = {k: v.to(dev) for k, v in bench.host_batch(n, 0, False).items()}
from upnerf_b200.utils import ray as ray_utils  # noqa: E402


def render():
    with torch.no_grad():
        o, d = ray_utils.get_rays(b["directions"], b["c2w"])
        rays = torch.cat([o, d, b["ray_infos"]], 1)
        return system(rays, b["feats"], b["img_idx"], 1.0, train=False)["rgb_fine"]


for _ in range(2):
    render()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    render()
    torch.cuda.synchronize()
with tempfile.TemporaryDirectory() as d:
    path = Path(d) / "trace.json"
    prof.export_chrome_trace(str(path))
    tr = json.loads(path.read_text())
evs = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
evs.sort(key=lambda e: e["ts"])
marks = [i for i, e in enumerate(evs) if "stratified_z" in e["name"]]
lo, hi = marks[2], marks[3]
step = evs[lo:hi]
t0 = step[0]["ts"]
streams = sorted({e["args"].get("stream", -1) for e in step})
last_end = {}
print(f"# one 4096-ray inference chunk: {len(step)} activities, span {step[-1]['ts'] + step[-1]['dur'] - t0:.1f} us; whole call {evs[-1]['ts'] + evs[-1]['dur'] - evs[0]['ts']:.1f} us for 6 chunks")
for e in step:
    s = e["args"].get("stream", -1)
    gap = e["ts"] - last_end.get(s, e["ts"])
    last_end[s] = e["ts"] + e["dur"]
    name = e["name"].replace("upnerf::(anonymous namespace)::", "").replace("void ", "")
    print(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} {gap:7.1f}  {streams.index(s):3d}    {name[:70]}")
