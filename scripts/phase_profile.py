"""Per-phase cycle breakdown of the step kernel (clock64 stamps, lane 0 of every warp)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tds_b200, tds_b200.workloads as wl
n = 4096
sim = tds_b200.laikago_sim(n)
w = wl.laikago(n)
sim.env_set_state(w["q"], w["qd"])
act = sim.alloc(12)
for _ in range(200):
    sim.env_step_device(act)
L = tds_b200.lib()
L.tds_b200_debug_phase_clocks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
nw = L.tds_b200_debug_phase_clocks(sim._h, 1, None, 0)
for _ in range(3):
    sim.env_step_device(act)
buf = np.zeros((nw, 16), dtype=np.int64)
L.tds_b200_debug_phase_clocks(sim._h, 1, buf.ctypes.data, nw)
print("kernel:", sim.kernel_name())
names = ["load+PD", "pass1 FK+contacts", "pass2 ABA+CRBA", "base+pass3", "cholesky", "J+Y", "PGS", "backsub", "integrate+write"]
K = len(names)
kern = os.environ.get("TDS_B200_KERNEL", "spec")
if kern in ("role", "spec"):   # one record per (tile, role): show every role, phases are separated by CTA barriers
    tiles = (n + 31) // 32
    b = buf[:tiles * 4].reshape(tiles, 4, 16)
    t0 = b[:, :, 0].min(axis=1, keepdims=True)
    tot = (b[:, :, K].max(axis=1) - t0[:, 0]).astype(np.float64)
    print("cycles per tile: total median %.0f (min %.0f max %.0f)" % (np.median(tot), tot.min(), tot.max()))
    print("  phase end (cycles since tile start), median over tiles:  role0 role1 role2 role3 | role-0 duration")
    prev = np.zeros(4)
    for k, nm in enumerate(names):
        end = np.median((b[:, :, k + 1] - t0).astype(np.float64), axis=0)
        print(f"  {nm:18s} " + " ".join(f"{x:8.0f}" for x in end) + f" | {end[0] - prev[0]:8.0f}")
        prev = end
    if kern == "spec":
        ex = np.median((b[:, 0, 10:14] - t0).astype(np.float64), axis=0)
        print("  role 0 extra stamps (cycles since tile start): pass1a start %.0f end %.0f | pass2a end %.0f | trunk leaf->root end %.0f" % (ex[2], ex[3], ex[0], ex[1]))
    sys.exit(0)

d = np.diff(buf[:, :K + 1], axis=1).astype(np.float64)
tot = (buf[:, K] - buf[:, 0]).astype(np.float64)
print("cycles per warp: total median %.0f (min %.0f max %.0f)" % (np.median(tot), tot.min(), tot.max()))
for k, nm in enumerate(names):
    print(f"  {nm:16s} {np.median(d[:, k]):9.0f} cycles  {100 * np.median(d[:, k]) / np.median(tot):5.1f}%")
