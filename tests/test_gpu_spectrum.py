"""
-m gpu parity tests of the P(k) step that follows BaryonifySnapshot.process() in the reference's workflow
(examples/10_Reproduce_Schneider_deltaPk.ipynb cells 1, 12, 15): folded particle deposit,
FFT, k-shell sums on the device, through the C ABI, against
  (1) the committed fixtures = output of the notebook's own cell code (tests/golden/pk_nb10_*.npz, oracle/make_golden.py),
  (2) the numpy restatement oracle/pk_port.py on fresh inputs.
Shell membership (mode counts) and the particle deposit are bit-exact; P(k) agrees to relative 1e-6 (cuFFT vs pocketfft).
"""
import numpy as np
import pytest

from helpers import assert_close, golden_names, load

pytestmark = pytest.mark.gpu


def _bin_half_spectrum(N, Nk, L, spec):
    """bfg_power_bin_spectrum on a host half spectrum (complex128 [N][N][N//2+1]); returns (pk_sum, k_sum, count)."""
    import torch
    from baryonforge_b200 import _lib
    from oracle import pk_port
    kb = pk_port.KBinning(N, Nk, L)
    dev = torch.device('cuda', torch.cuda.current_device())
    d_klin = torch.from_numpy(kb.klin.copy()).to(dev)
    d_spec = None if spec is None else torch.view_as_real(torch.from_numpy(np.ascontiguousarray(spec)).to(dev)).contiguous()
    out = torch.full((2, Nk), -1.0, dtype=torch.float64, device=dev)           # the call zeroes its outputs
    cnt = torch.full((Nk,), -1, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().bfg_power_bin_spectrum(N, _lib.ptr(d_spec), d_klin.data_ptr(), float(kb.kbins[0]),
                                                 float(kb.kbins[1] - kb.kbins[0]), Nk, out[0].data_ptr(),
                                                 out[1].data_ptr(), cnt.data_ptr(), _lib.current_stream()))
    o = out.cpu().numpy()
    return kb, o[0], o[1], cnt.cpu().numpy()


@pytest.mark.parametrize("N,Nk", [(16, 10), (27, 12), (48, 30), (9, 5), (66, 200), (128, 180)])
def test_shell_sums_of_a_half_spectrum_match_numpy(N, Nk):
    """Even, odd and non-multiple-of-32 row lengths; more shells than modes per row (empty shells)."""
    rng = np.random.default_rng(N)
    grid = rng.poisson(2.0, (N, N, N)).astype(np.float64)
    kb, pk_sum, k_sum, cnt = _bin_half_spectrum(N, Nk, 100.0, np.fft.rfftn(grid))
    assert np.array_equal(cnt, kb.k_c)                                         # shell membership bit-exact
    with np.errstate(invalid='ignore', divide='ignore'):
        assert np.allclose(k_sum / cnt, kb.k_cen, rtol=1e-12, atol=0, equal_nan=True)
        got, want = pk_sum / cnt, kb.power(grid)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.allclose(got[ok], want[ok], rtol=1e-10, atol=0)
    # mode counting only (NULL spectrum)
    _, pk0, k0, c0 = _bin_half_spectrum(N, Nk, 100.0, None)
    assert np.array_equal(c0, kb.k_c) and np.all(pk0 == 0) and np.allclose(k0, k_sum, rtol=1e-14)


@pytest.mark.parametrize("name", golden_names("pk"))
def test_pk_matches_notebook_fixture(name):
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import pk_port
    g = load(name)
    N, Nk, L = int(g["Ngrd"]), int(g["Nk"]), float(g["L"])
    sp = b.ShellPowerSpectrum(N, Nk, L)
    assert np.array_equal(sp.kbins, g["kbins"]) and np.array_equal(sp.klin, g["klin"])
    assert np.array_equal(sp.k_c, g["k_c"])
    assert np.allclose(sp.k_cen, g["k_cen"], rtol=1e-12, atol=0, equal_nan=True)
    p = synth.pk_particles(int(g["n_part"]), L, int(g["seed"]))
    for factor, key in ((1, "pk_f1"), (8, "pk_f8")):
        d_grid = sp.deposit_on_device(p, factor)
        want_grid = pk_port.histogram3d(pk_port.fold_positions(p, L / factor), N, 0.0, L / factor)
        assert np.array_equal(d_grid.cpu().numpy(), want_grid.astype(np.float64))    # particle assignment bit-exact
        assert sp.last_dropped == 0
        assert_close(sp.measure_grid(d_grid), g[key], f"{name} factor {factor}")     # rel 1e-6
        assert_close(sp.measure(p, factor), g[key], f"{name} factor {factor} (measure)")


def test_folded_deposit_edge_cases():
    """Negative and >= L coordinates fold like numpy's %, non-finite ones are dropped and counted, a folded value that
    rounds up to L_fold goes to the last cell, coordinates may come as device tensors."""
    import torch
    import baryonforge_b200 as b
    from oracle import pk_port
    N, L = 8, 4.0
    x = np.array([-0.5, 4.0, 9.0, -1e-20, 3.999999, 0.0, np.nan, 1.0, np.inf, 2.0])
    y = np.array([0.25, 0.75, 1.25, 1.75, 2.25, 2.75, 3.25, np.nan, 3.75, -4.0])
    z = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 8.5])
    sp = b.ShellPowerSpectrum(N, 4, L)
    got = sp.deposit_on_device([x, y, z], 1).cpu().numpy()
    assert sp.last_dropped == 3 and got.sum() == 7
    ok = np.isfinite(x) & np.isfinite(y) & np.isfinite(z)
    pts = np.stack([x[ok], y[ok], z[ok]], axis=1)
    folded = pk_port.fold_positions(pts, L)
    folded[folded >= L] = np.nextafter(L, 0)                  # -1e-20 % 4 == 4.0: out of bounds in the notebook's numba loop
    assert np.array_equal(got, pk_port.histogram3d(folded, N, 0.0, L).astype(np.float64))
    dev = torch.device('cuda', torch.cuda.current_device())
    d = [torch.from_numpy(a).to(dev) for a in (x, y, z)]
    assert np.array_equal(sp.deposit_on_device(d, 1).cpu().numpy(), got)
    for factor in (2, 4):
        g2 = sp.deposit_on_device(pts, factor).cpu().numpy()
        f2 = pk_port.fold_positions(pts, L / factor)
        f2[f2 >= L / factor] = np.nextafter(L / factor, 0)
        assert np.array_equal(g2, pk_port.histogram3d(f2, N, 0.0, L / factor).astype(np.float64))
    with pytest.raises(ValueError):
        sp.deposit_on_device([x, y], 1)


def test_snapshot_on_device_then_pk_matches_host_path():
    """BaryonifySnapshot.process_on_device() -> P(k) without the particles leaving HBM == process() -> oracle P(k)."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import pk_port
    g = load("snap_3d")
    cosmo = synth.COSMO
    mc = dict(Omega_m=0.27 + 0.05, Omega_b=0.05, h=0.68, sigma8=0.82, n_s=0.97, w0=-1.0)
    L = float(g["L"])
    cat = b.HaloNDCatalog(x=g["x"], y=g["y"], z=g["z"], M=g["M"], redshift=g["redshift"], cosmo=cosmo)
    ps = b.ParticleSnapshot(x=g["px"], y=g["py"], z=g["pz"], M=g["pM"], L=L, redshift=g["redshift"], cosmo=cosmo)
    model = b.DisplacementModel((g["ax0"], g["ax1"], g["ax2"]), g["values"], g["eps_mod"], mc)
    run = b.BaryonifySnapshot(cat, ps, g["eps_run"], model, verbose=False)
    d_p = run.process_on_device()
    out = run.process()
    for k, name in enumerate("xyz"):
        assert np.array_equal(d_p[k].cpu().numpy(), out[name], equal_nan=True)
    N, Nk = 32, 20
    sp = b.ShellPowerSpectrum(N, Nk, L)
    kb = pk_port.KBinning(N, Nk, L)
    pts = np.stack([g["out_x"], g["out_y"], g["out_z"]], axis=1)              # the reference's own displaced particles
    for factor in (1, 2):
        want, want_grid = pk_port.folded_power(pts, kb, factor)
        got_grid = sp.deposit_on_device(d_p, factor)
        # a particle within round-off of a cell edge may land on the other side (positions agree to 1e-12)
        assert np.abs(got_grid.cpu().numpy() - want_grid).sum() <= 4
        assert_close(sp.measure_grid(want_grid.astype(np.float64)), want, f"snap_3d P(k) factor {factor}")
        assert_close(sp.measure(d_p, factor), want, f"snap_3d P(k) on device, factor {factor}", rtol=1e-3)


def test_pk_from_the_cell_ordered_particles_equals_the_caller_ordered_path():
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    g = load("snap_3d")
    cosmo = synth.COSMO
    mc = dict(Omega_m=0.27 + 0.05, Omega_b=0.05, h=0.68, sigma8=0.82, n_s=0.97, w0=-1.0)
    L = float(g["L"])
    cat = b.HaloNDCatalog(x=g["x"], y=g["y"], z=g["z"], M=g["M"], redshift=g["redshift"], cosmo=cosmo)
    ps = b.ParticleSnapshot(x=g["px"], y=g["py"], z=g["pz"], M=g["pM"], L=L, redshift=g["redshift"], cosmo=cosmo)
    model = b.DisplacementModel((g["ax0"], g["ax1"], g["ax2"]), g["values"], g["eps_mod"], mc)
    run = b.BaryonifySnapshot(cat, ps, g["eps_run"], model, verbose=False)
    sp = b.ShellPowerSpectrum(32, 20, L)
    d_p = run.process_on_device()
    want = [sp.measure(d_p, f) for f in (1, 2, 8)]
    got = sp.measure_runner(run, factors=(1, 2, 8))
    for a, w in zip(got, want):
        assert_close(a, w, "P(k) from the cell-ordered particles", rtol=1e-9)      # same grid, FFT round-off only
    assert run.last_stats["n_pairs"] > 0 and sp.last_dropped == 0

