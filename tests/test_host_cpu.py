"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol the header
declares, the custom ops are registered for CUDA only, and argument errors surface as exceptions."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import stereoanywhere_b200._lib as L

    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sa_b200.h but not exported"
    assert sorted(L.EXPORTS) == names
    assert lib.sa_abi_version() == 1


def test_library_is_in_tree_and_self_contained():
    import stereoanywhere_b200._lib as L

    assert L.LIB_PATH.startswith(ROOT) and os.path.exists(L.LIB_PATH)
    # no torch / libcudart.so runtime dependency: plain C ABI, static cudart
    import subprocess

    deps = subprocess.run(["ldd", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in deps and "libc10" not in deps


def test_argument_errors_return_negative_codes_without_a_gpu():
    import stereoanywhere_b200._lib as L

    lib = L.load()
    rc = lib.sa_corr_fp32(None, None, None, 1, 3, 4, 8, 8, 1.0, 1.0, None)
    assert rc == -1 and b"null" in lib.sa_last_error()
    rc = lib.sa_pyramid(None, 0, 8, 8, 1, None, None, None, 0, 0, 0, None, None, 0.0, 0, None, None)
    assert rc == -1
    rc = lib.sa_lookup(None, None, None, 4, 4, None, 0, None, 1, 1, 8, 0, 0, None)
    assert rc == -1
    with pytest.raises(L.SaError):
        L.check(rc, "sa_lookup")


def test_ops_registered_cuda_only():
    import stereoanywhere_b200 as sa

    for name in sa.ops.OP_NAMES:
        assert hasattr(torch.ops.sa_b200, name)
    f = torch.zeros(1, 8, 2, 8)
    with pytest.raises(NotImplementedError):
        sa.CorrBlockB200.corr(f, f)  # CPU tensors: no kernel registered, no fallback
    with pytest.raises(NotImplementedError):
        sa.CorrBlockB200(torch.zeros(1, 2, 8, 1, 8))


def test_level_geometry():
    from stereoanywhere_b200 import ops

    assert ops.level_widths(312, 4) == [312, 156, 78, 39]
    assert ops.level_widths(39, 4) == [39, 19, 9, 4]
    assert [ops.level_pitch(w) for w in (312, 156, 78, 39)] == [312, 156, 80, 40]


def test_grad_policy():
    """Gradients flow to volumes / feature maps; coords and the truncation maps must be detached."""
    from stereoanywhere_b200 import corr as C

    t = torch.zeros(1, requires_grad=True)
    assert C._needs_grad(t) and not C._needs_grad(t.detach())
    with torch.no_grad():
        assert not C._needs_grad(t)
    with pytest.raises(NotImplementedError):
        C._no_grad_check(None, t)
    C._no_grad_check(None, t.detach())


def test_packed_row_floats_matches_library():
    from stereoanywhere_b200 import _lib, ops

    lib = _lib.load()
    for w in (8, 40, 128, 312, 768, 1024):
        assert ops.packed_row_floats(w) == int(lib.sa_packed_row_floats(w))


def test_sass_uses_the_blackwell_units():
    """The shipped library is sm_100a code that really issues tcgen05 MMAs out of TMEM and TMA loads / stores
    (SASS mnemonics of /opt/skills/guides/B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG /
    UTMASTG (cp.async.bulk.tensor), LDGSTS (cp.async) - not a recompiled mma.sync / SIMT fallback."""
    import shutil
    import subprocess

    import stereoanywhere_b200._lib as L

    L.load()
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", L.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    for mnemonic, least in (("UTCHMMA", 8), ("LDTM", 3), ("UTMALDG", 8), ("UTMASTG", 4), ("LDGSTS", 8)):
        assert sass.count(mnemonic) >= least, f"{mnemonic}: {sass.count(mnemonic)} occurrences"
    assert "HMMA.16816" not in sass and "WGMMA" not in sass


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract's keys, whatever libraries write to file descriptor 1."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1_384x512_b1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["unit"] == "pairs/s" and out["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    # the reference's own CorrBlock1D is timed whenever its files are there (here: /root/reference or oracle/_ref)
    from oracle import ref_path

    assert out["cpu_baseline"]["kind"] == ("reference" if ref_path.available() else "port")
    assert out["cpu_baseline"]["cores"] >= 1 and out["warmup"] == 0
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["e2e"]["d2h_bytes_per_step"] == 0
    # both arms print the same `config` (the driver compares them)
    import bench

    assert out["config"] == bench.workload_config("c1_384x512_b1", 1)


def test_oracle_ref_recipe_and_reference_path():
    """`oracle/make_ref.py` reproduces the reference files byte for byte (the copy that travels to the GPU box), and
    the path run by the reference's own classes equals the oracle's restatement bit for bit."""
    import torch

    from oracle import corr_oracle as O
    from oracle import make_ref, ref_path

    if not ref_path.available():
        pytest.skip("no reference files in this environment")
    if os.path.isdir("/root/reference/models/stereoanywhere"):
        assert make_ref.make() and make_ref.verify()
        import json

        files = json.load(open(os.path.join(make_ref.DEST, "MANIFEST.json")))["files"]
        assert "models/stereoanywhere/corr.py" in files and "mapreduce_v2/tile_wrapper.py" in files
        for rel in files:
            assert open(os.path.join("/root/reference", rel), "rb").read() == open(os.path.join(make_ref.DEST, rel), "rb").read()
    g = torch.Generator().manual_seed(11)
    b, c, h, w = 1, 16, 3, 40
    fl, fr = torch.randn(b, c, h, w, generator=g), torch.randn(b, c, h, w, generator=g)
    nl = torch.nn.functional.normalize(torch.randn(b, 3, h, w, generator=g), dim=1)
    nr = torch.nn.functional.normalize(torch.randn(b, 3, h, w, generator=g), dim=1)
    x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(b, 1, h, w)
    coords = [torch.cat([x - torch.rand(b, 1, h, w, generator=g) * 10, torch.zeros(b, 1, h, w)], 1) for _ in range(2)]
    trunc = (torch.rand(b, 1, h, w, generator=g) * 10, torch.rand(b, 1, h, w, generator=g), 0.9)
    rs, rm = ref_path.run_path_reference(fl, fr, nl, nr, coords, trunc=trunc)
    ps, pm = O.run_path_cpu(fl, fr, nl, nr, coords, trunc=trunc)
    assert torch.equal(rs, ps) and torch.equal(rm, pm)


def test_integration_install_uninstall_swaps_the_reference_names():
    """`integration.install` only rebinds two names of the reference module and `uninstall` restores them."""
    import types

    from stereoanywhere_b200 import CorrBlockB200, integration

    mod = types.SimpleNamespace(CorrBlock1D=object(), truncate_corr_volume_v2=lambda *a, **k: "ref")
    orig = (mod.CorrBlock1D, mod.truncate_corr_volume_v2)
    integration.install(mod)
    assert mod.CorrBlock1D is CorrBlockB200 and mod.truncate_corr_volume_v2 is orig[1]
    integration.install(mod, fused=True)
    assert mod.CorrBlock1D is integration.FusedCorrBlock and mod.truncate_corr_volume_v2 is not orig[1]
    import torch

    d = torch.zeros(1, 1, 2, 8)
    assert mod.truncate_corr_volume_v2(d, d, conf_th=None, attenuation_gain=0.9) == "ref"   # CPU tensors: reference function
    integration.uninstall(mod)
    assert (mod.CorrBlock1D, mod.truncate_corr_volume_v2) == orig
    # the symbolic reshapes of stereoanywhere.py:135-136, 253-258
    lv = integration.LazyVolume(torch.zeros(2, 3, 4, 8), torch.zeros(2, 3, 4, 16))
    assert lv.shape == [2, 4, 8, 1, 16]
    v = (1.73 * lv.squeeze(3).unsqueeze(1))
    assert v.shape == [2, 1, 4, 8, 16] and abs(v.gain - 1.73) < 1e-12
    assert v.squeeze(1).unsqueeze(3).shape == [2, 4, 8, 1, 16]
