#!/usr/bin/env python
"""tools/bench_anafast.py -- device timing of the C_l step (harmonics.ShellHarmonics, csrc/sht_kernels.cu), CUDA events.
One JSON line: per NSIDE the time of one analysis pass, one synthesis and a full anafast (iter = 3: seven transforms)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def events(fn, warm=1, reps=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsides", type=int, nargs="+", default=[256, 512, 1024])
    args = ap.parse_args()
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    out = []
    for nside in args.nsides:
        sh = b.harmonics.ShellHarmonics(nside)
        d_map = torch.rand(sh.npix, dtype=torch.float64, device=dev) * 10
        d_ln, d_work = sh._buffers(dev)
        d_alm = torch.zeros((sh.n_alm, 2), dtype=torch.float64, device=dev)
        d_syn = torch.empty_like(d_map)
        st = torch.cuda.current_stream().cuda_stream

        def ana():
            d_alm.zero_()
            _lib.check(L.bfg_sht_map2alm_pass(nside, sh.lmax, d_map.data_ptr(), d_ln.data_ptr(), d_work.data_ptr(),
                                              d_alm.data_ptr(), st))

        def syn():
            _lib.check(L.bfg_sht_alm2map(nside, sh.lmax, d_alm.data_ptr(), d_ln.data_ptr(), d_work.data_ptr(), d_syn.data_ptr(), st))
        case = dict(nside=nside, lmax=sh.lmax, analysis_ms=events(ana), synthesis_ms=events(syn))
        case["anafast_iter3_ms"] = events(lambda: sh.alm2cl_on_device(sh.map2alm_on_device(d_map, 3)), warm=0, reps=1)
        cl = sh.alm2cl_on_device(sh.map2alm_on_device(d_map, 0)).cpu().numpy()
        ells = np.arange(sh.lmax + 1)
        case["parseval_ratio_iter0"] = float(np.sum((2 * ells + 1) * cl) / (4 * np.pi / sh.npix * float((d_map * d_map).sum())))
        out.append(case)
    print(json.dumps(dict(cases=out)))


if __name__ == "__main__":
    main()
