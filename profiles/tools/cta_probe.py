import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np
from pdffoam_b200 import cases
from pdffoam_b200.capi import DeviceLibrary
lib = DeviceLibrary()
case = cases.synthetic_box(200, 100, 100, ppc=64, closed=False, number_control=True)
h = case.make_cloud(lib, device=0, capacity=int(case.mesh.n_cells * 64 * 1.2) + 65536)
h.seed_particles()
h.profile_ctas()
for target in (23, 405):
    h.evolve(target - int(h.step_counter()) if hasattr(h, "step_counter") else target)
    h.evolve(1)
    t = h.profile_ctas()
    dur = (t[:, 1] - t[:, 0]) * 1e-6
    start = (t[:, 0] - t[:, 0].min()) * 1e-6
    end = (t[:, 1] - t[:, 0].min()) * 1e-6
    o = np.argsort(-end)
    print("after step", target, "ctas", len(t), "dur ms: min %.2f median %.2f p90 %.2f p99 %.2f max %.2f; kernel span %.2f" % (dur.min(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), end.max()))
    print("  latest CTAs (index, sm, start, end):", [(int(i), int(t[i, 2]), round(float(start[i]), 2), round(float(end[i]), 2)) for i in o[:12]])
    sm_end = {}
    for i in range(len(t)): sm_end[int(t[i, 2])] = max(sm_end.get(int(t[i, 2]), 0), end[i])
    e = np.array(sorted(sm_end.values()))
    print("  per-SM end ms: min %.2f median %.2f max %.2f; SMs later than median+1ms: %d" % (e.min(), np.median(e), e.max(), int((e > np.median(e) + 1).sum())))
    hist, edges = np.histogram(end, bins=12)
    print("  end-time histogram:", list(zip([round(float(x), 1) for x in edges[:-1]], hist.tolist())))
    np.save("gpurun_out/g29_cta_%d.npy" % target, t)
