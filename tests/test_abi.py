"""The C-ABI library loads on a CPU-only box and exports every symbol include/tensorf_b200.h
declares (no compute calls without a GPU); argument validation that needs no device."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "tensorf_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tensorf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from tensorf_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 16
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    for n in names:
        assert hasattr(lib, n), f"{n} not exported"
    assert lib.tensorf_version() >= 100


def test_struct_layouts_match_header():
    from tensorf_b200 import _lib
    assert ctypes.sizeof(_lib.RenderDesc) == 16 * 4
    assert ctypes.sizeof(_lib.Params) == 12 * 8
    assert ctypes.sizeof(_lib.RenderInputs) == 9 * 8


def test_host_side_validation_without_gpu():
    from tensorf_b200 import _lib, ops
    lib = _lib.load()
    assert lib.tensorf_vm_packed_floats(16, 128) == 3 * 128 * 16 + 3 * 128 * 128 * 16
    assert lib.tensorf_vm_packed_floats(5, 4) == 3 * 4 * 8 + 3 * 16 * 8          # channels padded to 8
    nbytes = ctypes.c_int64()
    bad = ops.make_desc(R=4, N=5, K=6, G=9, cd=2, ca=3)                            # K > N
    assert lib.tensorf_render_workspace_bytes(ctypes.byref(bad), ctypes.byref(nbytes)) == -1
    assert b"appearance_samples_per_ray" in lib.tensorf_last_error()
    ok = ops.make_desc(R=4096, N=221, K=33, G=128, cd=16, ca=48, feat_freqs=2, view_freqs=2)
    assert lib.tensorf_render_workspace_bytes(ctypes.byref(ok), ctypes.byref(nbytes)) == 0 and nbytes.value > 0
    un = ops.make_desc(R=4, N=5, K=2, G=9, cd=2, ca=3, units=64)
    assert lib.tensorf_render_workspace_bytes(ctypes.byref(un), ctypes.byref(nbytes)) == -3
    with pytest.raises(_lib.TensorfError):
        _lib.check(-1)


def test_missing_library_fails_loudly(tmp_path):
    from tensorf_b200 import _lib
    with pytest.raises(_lib.TensorfLibraryError, match="no CPU fallback"):
        _lib.load(tmp_path / "libtensorf_b200.so")


def test_ops_reject_cpu_tensors():
    import torch
    from tensorf_b200 import ops
    with pytest.raises(ValueError, match="no CPU path"):
        ops._ptr(torch.zeros(3))


def test_library_threefry_known_answers():
    """The host cipher block exported by the library (no GPU needed) against the Random123 vectors."""
    from tensorf_b200 import ops
    kat = [((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, out in kat:
        assert ops.threefry2x32(key[0], key[1], ctr[0], ctr[1]) == out


def test_peer_shard_ranges_and_host_validation():
    """tensorf_peer_shard (host function) tiles [0,total) with 4-aligned ranges; tensorf_adam_step_peer validates
    its descriptor before touching the device."""
    from tensorf_b200 import _lib
    lib = _lib.load()
    for total in (0, 4, 8, 100, 3210420):
        for world in (1, 2, 3, 8, 16):
            prev, sizes = 0, []
            for r in range(world):
                b, e = ctypes.c_int64(), ctypes.c_int64()
                lib.tensorf_peer_shard(total, r, world, ctypes.byref(b), ctypes.byref(e))
                assert b.value == prev and b.value % 4 == 0 and e.value % 4 == 0 and e.value >= b.value
                prev = e.value
                sizes.append(e.value - b.value)
            assert prev == total and max(sizes) - min(sizes) <= 4
    assert ctypes.sizeof(_lib.PeerAdamDesc) == 72
    assert lib.tensorf_peer_adam_scratch_bytes(0) >= 16 and lib.tensorf_peer_adam_scratch_bytes(-1) == -1
    fake = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    offs = (ctypes.c_int64 * 2)(0, 8)
    neg = (ctypes.c_float * 1)(-1.0)

    def call(**kw):
        f = dict(rank=0, world=2, total=8, shard_begin=0, shard_end=4, n_leaves=1)
        f.update(kw)
        d = _lib.PeerAdamDesc(adam=_lib.AdamDesc(n_leaves=f["n_leaves"], reserved=0, b1=0.9, b2=0.99, eps=1e-8, eps_root=0.0,
                                                 bias_correction1=0.1, bias_correction2=0.01, lr_decay=1.0, reserved2=0.0),
                              rank=f["rank"], world=f["world"], total=f["total"], shard_begin=f["shard_begin"],
                              shard_end=f["shard_end"])
        return lib.tensorf_adam_step_peer(None, ctypes.byref(d), offs, neg, fake, fake, None, None, 0x3000, 0x4000, fake,
                                          None, 0)
    assert call() == -1 and b"scratch" in lib.tensorf_last_error()            # everything else valid: stops at scratch
    assert call(world=0) == -1 and b"world" in lib.tensorf_last_error()
    assert call(rank=2) == -1 and b"rank" in lib.tensorf_last_error()
    assert call(total=10) == -1 and b"multiple of 4" in lib.tensorf_last_error()
    assert call(shard_end=6) == -1 and b"shard" in lib.tensorf_last_error()
    assert call(n_leaves=33) == -1 and b"n_leaves" in lib.tensorf_last_error()
