import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_step_graph_gpu import _run
dev = torch.device("cuda:0")
def diff(a, b, sd):
    out = {}
    for k, v in a.items():
        if k.endswith("progress"): continue
        ue, ug = v - sd[k], b[k] - sd[k]
        n = float(ue.norm())
        out[k] = float((ue - ug).norm()) / n if n else float(ug.norm())
    return out
for prec, start in (("fp32", 0.29),):
    e1 = _run(dev, False, prec, 0.0, 9, start)
    e2 = _run(dev, False, prec, 0.0, 9, start)
    g1 = _run(dev, True, prec, 0.0, 9, start)
    d_ee = diff(e1[3], e2[3], e1[4]); d_eg = diff(e1[3], g1[3], e1[4])
    print("eager vs eager worst", max(d_ee.values()), "eager vs graph worst", max(d_eg.values()))
    for k in sorted(d_eg, key=lambda k: -d_eg[k])[:12]:
        print(f"  {k:50s} ee {d_ee[k]:.3e}  eg {d_eg[k]:.3e}")
    print("losses e", [round(x, 6) for x in e1[1]]); print("losses g", [round(x, 6) for x in g1[1]])
    for n in (4, 5, 6, 7, 8):
        e = _run(dev, False, prec, 0.0, n, start); g = _run(dev, True, prec, 0.0, n, start)
        d = diff(e[3], g[3], e[4]); k = max(d, key=d.get)
        print(n, "steps: worst", k, d[k])
