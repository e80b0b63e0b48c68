"""Phase timing of BaryonifyShell.process() at the bench workload (BFG_PROFILE_E2E=1 adds syncs between phases)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import baryonforge_b200 as b
from baryonforge_b200 import synth
nside, n = 4096, 1000000
ra, dec, M, z = synth.sky_halos(n, seed=42)
axes = synth.table_axes()
cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
model = b.DisplacementModel(axes, synth.displacement_values(axes), 20, synth.COSMO)
pm = torch.empty(12 * nside * nside, dtype=torch.float64, pin_memory=True)
pm.numpy()[:] = synth.shell_map(nside, seed=7)
shell = b.LightconeShell(map=pm.numpy(), cosmo=synth.COSMO)
run = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
for mode in ("drop", "del", "hold", "profile-del"):
    os.environ["BFG_PROFILE_E2E"] = "1" if mode.startswith("profile") else "0"
    held = None
    for it in range(4):
        t0 = time.perf_counter()
        out = run.process()
        t1 = time.perf_counter()
        if mode == "hold":
            held = out
        elif mode in ("del", "profile-del"):
            del out
        else:
            out = None
        t2 = time.perf_counter()
        print(mode, it, "process_ms", round(1e3 * (t1 - t0), 1), "release_ms", round(1e3 * (t2 - t1), 1),
              json.dumps({k: round(1e3 * v, 1) for k, v in run.last_timing.items()}))
    held = None
