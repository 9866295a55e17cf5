"""Summarise gpurun_out/trace_w1.npy (tools/trace_w1.py): mean clocks each role spends waiting / working per step."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
t = np.load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "trace_w1.npy"))
names = ["wprod", "xprod", "mma0", "epi", "bld0", "bld1", "conv0", "conv1", "mma1"]
def seq(r):
    row = t[r]; row = row[row != 0]
    return [(int(v) >> 4, int(v) & 15) for v in row]
def spans(r, a, b, skip=100):
    """durations from event a to the next event b of role r"""
    s = seq(r)[skip:]; out = []; last = None
    for c, e in s:
        if e == b and last is not None and a != b:
            out.append(c - last); last = None
        if e == a: last = c
        elif False: out.append(c - last); last = None
    return np.array(out) if out else np.array([0])
m = seq(2); starts = [c for c, e in m if e == 4]
print("tile period", np.mean(np.diff(starts)[2:]))
for r, pairs in [(1, [(0, 1, "wait x_empty")]), (2, [(0, 1, "wait x_empty + x_full"), (1, 2, "wait a_full"), (2, 3, "issue 9 MMAs + commits"), (4, 5, "wait t_empty")]), (8, [(0, 1, "wait x_empty + x_full"), (1, 2, "wait a_full"), (2, 3, "issue 9 MMAs + commits"), (4, 5, "wait t_empty")]),
                 (3, [(0, 1, "wait t_full"), (1, 2, "drain accumulator")]), (4, [(0, 1, "wait w_full"), (2, 3, "wait a_empty"), (3, 4, "build chunk")]),
                 (5, [(0, 1, "wait w_full"), (2, 3, "wait a_empty"), (3, 4, "build chunk")]), (6, [(0, 1, "wait x_raw"), (1, 2, "convert stage")]), (7, [(0, 1, "wait x_raw"), (1, 2, "convert stage")]),
                 (0, [(0, 1, "wait w_empty")])]:
    for a, b, what in pairs:
        d = spans(r, a, b)
        print(f"{names[r]:6s} {what:28s} n={len(d):5d} mean {d.mean():8.0f} p50 {np.median(d):8.0f} max {d.max():8.0f}")
