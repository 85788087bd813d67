"""Two eager train steps of the bench workload (for ncu: profile kernels of the second step)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
system = bench.make_system("bf16", dev)
system.hparams["kernel.cuda_graph"] = False
b = {k: v.to(dev) for k, v in bench.host_batch(4096, 0, False).items()}
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    system.training_step(b, i)
torch.cuda.synchronize()
