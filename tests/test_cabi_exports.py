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
