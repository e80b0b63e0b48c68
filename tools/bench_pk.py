#!/usr/bin/env python
"""
tools/bench_pk.py -- device-resident timing of the P(k) step that follows BaryonifySnapshot (spectra.ShellPowerSpectrum;
examples/10_Reproduce_Schneider_deltaPk.ipynb cells 1, 12, 15), 1 GPU, CUDA events, for profiles/.

  deposit   bfg_snap_deposit_folded   24 B position read + 16 B f64 RMW = 40 B / particle
  fft+bins  bfg_grid_power_spectrum   cuFFT D2Z (library) + k_power_bins: 16 B / mode of the half spectrum

Prints one JSON line.  Particles are generated on the device (torch.rand = plumbing).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def events(fn, warm=1, reps=3):
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
    ap.add_argument("--n-part", type=float, default=2.5e8)
    ap.add_argument("--grids", type=int, nargs="+", default=[256, 512, 1024])
    ap.add_argument("--peak", type=float, default=6650.0, help="HBM GB/s (B200_PROFILING.md fallback)")
    args = ap.parse_args()
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    n = int(args.n_part)
    Lbox = 1000.0
    gen = torch.Generator(device=dev).manual_seed(1)
    d_p = [torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * Lbox for _ in range(3)]
    st = torch.cuda.current_stream().cuda_stream
    out = dict(n_part=n, Lbox=Lbox, peak_GBps=args.peak, cases=[])
    for N in args.grids:
        Nk = min(180 * N // 256, 4096)
        sp = b.ShellPowerSpectrum(N, Nk, Lbox)
        d_grid = torch.zeros((N,) * 3, dtype=torch.float64, device=dev)
        d_klin = torch.from_numpy(sp.klin.copy()).to(dev)
        o = torch.empty((2, Nk), dtype=torch.float64, device=dev)
        cnt = torch.empty(Nk, dtype=torch.int64, device=dev)
        k0, dk = float(sp.kbins[0]), float(sp.kbins[1] - sp.kbins[0])
        case = dict(Ngrd=N, Nk=Nk)
        for factor in (1, 8):
            def dep():
                d_grid.zero_()
                _lib.check(L.bfg_snap_deposit_folded(n, d_p[0].data_ptr(), d_p[1].data_ptr(), d_p[2].data_ptr(), Lbox / factor,
                                                     N, d_grid.data_ptr(), None, st))
            ms = events(dep)
            case["deposit_f%d_ms" % factor] = ms
            case["deposit_f%d_alg_GBps" % factor] = 40.0 * n / ms / 1e6
        assert float(d_grid.sum()) == n

        def spec():
            _lib.check(L.bfg_grid_power_spectrum(N, d_grid.data_ptr(), d_klin.data_ptr(), k0, dk, Nk, o[0].data_ptr(),
                                                 o[1].data_ptr(), cnt.data_ptr(), st))
        case["fft_bins_ms"] = events(spec)

        def bins_only():
            _lib.check(L.bfg_power_bin_spectrum(N, None, d_klin.data_ptr(), k0, dk, Nk, o[0].data_ptr(), o[1].data_ptr(),
                                                cnt.data_ptr(), st))
        case["bins_only_no_spectrum_ms"] = events(bins_only)
        case["modes_counted"] = int(cnt.sum())
        case["half_spectrum_GB"] = 16.0 * N * N * (N // 2 + 1) / 1e9
        out["cases"].append(case)
        del d_grid
    print(json.dumps(out))


if __name__ == "__main__":
    main()
