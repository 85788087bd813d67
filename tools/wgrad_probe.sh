# per-launch device times of the weight-gradient kernels at several M (ncu, serialised)
cd $GRAFT_REPO_ROOT
for m in 65536 262144 524288; do
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wgrad --csv \
    --log-file gpurun_out/wg_$m.csv python tools/bench_gemm.py M=$m wgrad > /dev/null 2>&1
  python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/wg_$m.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); gi = h.index("Grid Size")
seq = [(r[ki][:40], r[gi], float(r[vi].replace(",", ""))) for r in rows[1:]]
# 4 shapes x (13 atomic + 13 det pairs); print the median per (kernel, grid) in order of first appearance
by = collections.OrderedDict()
for i, (k, g, v) in enumerate(seq):
    by.setdefault((k, g, i // 13 if False else 0), []).append(v)
print("M=$m")
cur = None; acc = []
for k, g, v in seq + [("", "", 0)]:
    if (k, g) != cur:
        if acc:
            acc.sort(); print(f"  {cur[0]:42s} grid {cur[1]:14s} n={len(acc):3d} median {acc[len(acc)//2]/1e3:8.2f} us")
        cur, acc = (k, g), []
    acc.append(v)
PY
done
