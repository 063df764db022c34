"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports
every symbol include/nerf_b200.h declares; the Python surface mirrors the
reference's signatures; the product fails loudly without CUDA."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nerfb_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from nerf_projects_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from nerf_projects_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nerf_b200.h but not exported"
    # the ctypes table covers the header exactly
    assert sorted(_lib.exported_symbols()) == declared
    assert _lib.load().nerfb_abi_version() == 1


def test_python_surface_matches_reference_signatures():
    import nerf_projects_b200 as nb
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(nb.render) == ["H", "W", "K", "chunk", "rays", "c2w", "ndc", "near", "far", "use_viewdirs",
                              "c2w_staticcam", "kwargs"]
    assert sig(nb.render_rays) == ["ray_batch", "network_fn", "network_query_fn", "N_samples", "retraw", "lindisp",
                                   "perturb", "N_importance", "network_fine", "white_bkgd", "raw_noise_std",
                                   "verbose", "pytest"]
    assert sig(nb.batchify_rays) == ["rays_flat", "chunk", "kwargs"]
    assert sig(nb.run_network) == ["inputs", "viewdirs", "fn", "embed_fn", "embeddirs_fn", "netchunk"]
    assert sig(nb.raw2outputs) == ["raw", "z_vals", "rays_d", "raw_noise_std", "white_bkgd", "pytest"]
    assert sig(nb.sample_pdf) == ["bins", "weights", "N_samples", "det", "pytest"]
    assert sig(nb.get_rays) == ["H", "W", "K", "c2w"]
    assert sig(nb.ndc_rays) == ["H", "W", "focal", "near", "rays_o", "rays_d"]
    d = inspect.signature(nb.render).parameters
    assert d["chunk"].default == 1024 * 32 and d["ndc"].default is True and d["near"].default == 0. \
        and d["far"].default == 1.
    assert inspect.signature(nb.run_network).parameters["netchunk"].default == 1024 * 64


def test_state_dict_layout_matches_reference_module():
    import nerf_projects_b200 as nb
    from oracle import nerf_oracle as O
    for use_vd, icv, och in ((True, 27, 5), (False, 0, 4)):
        m = nb.NeRF(8, 256, 63, icv, och, [4], use_vd)
        ref = O.make_params(0, use_viewdirs=use_vd, input_ch_views=icv, output_ch=och)
        sd = m.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k in sd:
            assert tuple(sd[k].shape) == ref[k].shape, k
        m.load_state_dict(O.to_torch(ref))
    assert sum(p.numel() for p in nb.NeRF(8, 256, 63, 27, 5, [4], True).parameters()) == 595844
    e, d = nb.get_embedder(10, 0)
    assert d == 63 and nb.get_embedder(4, 0)[1] == 27 and nb.get_embedder(10, -1)[1] == 3


def test_no_cpu_fallback():
    """Ops refuse CPU tensors instead of silently computing somewhere else."""
    import nerf_projects_b200 as nb
    if torch.cuda.is_available():
        pytest.skip("checks the CPU-only behaviour")
    with pytest.raises((nb.NerfB200Error, RuntimeError, AssertionError)):
        nb.sample_pdf(torch.rand(4, 63).sort(-1)[0], torch.rand(4, 62), 16, det=True)
    with pytest.raises((nb.NerfB200Error, RuntimeError, AssertionError)):
        nb.raw2outputs(torch.rand(2, 8, 4), torch.rand(2, 8).sort(-1)[0], torch.rand(2, 3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nerf_projects_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|from\s+oracle|import_module\(.oracle", src, re.M), \
                f"{fn} imports the oracle"


def test_header_is_plain_c(tmp_path):
    """include/nerf_b200.h is the C ABI: it must compile as C99 (and as C++) on its own -- no torch / CUDA types."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "nerf_b200.h"\nint main(void) { return nerfb_abi_version() == NERFB_ABI_VERSION ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only"],
                ["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-x", "c++"]):
        res = subprocess.run(cmd + ["-I", inc, str(src)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_library_sass_uses_the_blackwell_tensor_and_copy_engines():
    """Static evidence that the shipped library is the sm_100a path (no mma.sync / wmma recompile): tcgen05.mma (UTCHMMA),
    tcgen05.ld (LDTM), tcgen05.commit (UTCBAR), tensor-map TMA loads (UTMALDG) and bulk async copies (UBLKCP) in its SASS."""
    import shutil
    import subprocess
    from nerf_projects_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_lib.LIB_PATH):
        from nerf_projects_b200 import build
        build.build()
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for op, at_least in (("UTCHMMA", 50), ("LDTM", 10), ("UTCBAR", 5), ("UTMALDG", 2), ("UBLKCP", 5)):
        assert sass.count(op) >= at_least, f"{op}: {sass.count(op)} occurrences"
    assert "HMMA.16816" not in sass and "WGMMA" not in sass          # no legacy warp-level tensor-core path


def test_precision_override_is_scoped_and_thread_local():
    """`with precision(...)` overrides the process default for the current thread only (process-global state)."""
    import threading

    import nerf_projects_b200 as nb
    base = nb.get_precision()
    seen = {}
    with nb.precision("fp32"):
        assert nb.get_precision() == "fp32"
        t = threading.Thread(target=lambda: seen.setdefault("other", nb.get_precision()))
        t.start(); t.join()
        with nb.precision("bf16"):
            assert nb.get_precision() == "bf16"
        assert nb.get_precision() == "fp32"
    assert nb.get_precision() == base and seen["other"] == base
    with pytest.raises(ValueError):
        with nb.precision("fp16"):
            pass
