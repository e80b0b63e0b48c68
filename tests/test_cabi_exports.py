"""CPU: the C-ABI library loads and exports every symbol include/bfg_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bfg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bfg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("bfg_table_create", "bfg_shell_offsets", "bfg_shell_paint", "bfg_shell_regrid", "bfg_grid_offsets",
                 "bfg_grid_paint", "bfg_grid_regrid", "bfg_snap_build_cells", "bfg_snap_offsets", "bfg_snap_apply",
                 "bfg_snap_deposit_ngp", "bfg_healpix_query_disc", "bfg_shell_baryonify_host"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from baryonforge_b200 import _build, _lib
    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    lib = ctypes.CDLL(_build.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/bfg_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.exported_symbols()) == declared_symbols()
    assert _lib.lib().bfg_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    ra, dec, M, z = synth.sky_halos(10)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.ones(12 * 16 * 16), cosmo=synth.COSMO)
    axes = synth.table_axes()
    model = b.DisplacementModel(axes, synth.displacement_values(axes), 20, synth.COSMO)
    with pytest.raises(b._lib.BFGError):
        b.BaryonifyShell(cat, shell, 20, model, verbose=False).process()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "baryonforge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_cufft_resolves_at_first_use_and_pk_has_no_cpu_fallback():
    """bfg_grid_power_spectrum loads cuFFT with dlopen (no load-time dependency of libbfg_b200.so): without a GPU the call
    must get past the loader (status BFG_ERR_CUDA = -2 from the runtime, not BFG_ERR_UNSUPPORTED = -3), and the Python
    front end must raise instead of computing anything on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import baryonforge_b200 as b
    L = b._lib.lib()
    a, c = np.zeros(8), np.zeros(8, dtype=np.int64)
    rc = L.bfg_grid_power_spectrum(2, a.ctypes.data, a.ctypes.data, 1.0, 1.0, 4, a.ctypes.data, a.ctypes.data, c.ctypes.data,
                                   None)
    assert rc == -2, (rc, L.bfg_last_error())
    sp = b.ShellPowerSpectrum(16, 8, 100.0)
    assert sp.kbins.size == 9 and sp.klin.size == 16
    with pytest.raises(b._lib.BFGError):
        sp.measure(np.zeros((4, 3)))
    with pytest.raises(b._lib.BFGError):
        sp.k_c
