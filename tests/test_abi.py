"""The C-ABI library loads and exports every symbol include/gsv_cuda.h declares; without a GPU the
compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gsv_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsv_[a-z0-9_]+)\s*\(", src)) - {"gsv_body_fn"})


def test_all_header_symbols_exported(gsv):
    lib = gsv.load_library()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libgsv_cuda.so does not export {n}"


def test_product_does_not_link_oracle(gsv):
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", gsv.LIB_PATH], capture_output=True, text=True).stdout
    assert "gsvo_" not in out
    deps = subprocess.run(["ldd", gsv.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_record_api_from_c_callbacks(gsv):
    """gsv_program_record + gsv_ctx_*: a half adder described through the C callback ABI."""
    lib = gsv.load_library()
    BODY = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32,
                            ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32)
    lib.gsv_ctx_issue_wire.restype = ctypes.c_uint32
    lib.gsv_ctx_issue_wire.argtypes = [ctypes.c_void_p]
    lib.gsv_ctx_add_gate.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_uint32] * 3
    lib.gsv_ctx_component.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32,
                                      ctypes.c_uint32, BODY, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32)]
    lib.gsv_program_record.restype = ctypes.c_void_p
    lib.gsv_program_record.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, BODY, ctypes.c_void_p, ctypes.c_void_p]

    @BODY
    def half_adder(ctx, user, ins, n_in, outs, arity):
        s, c = lib.gsv_ctx_issue_wire(ctx), lib.gsv_ctx_issue_wire(ctx)
        lib.gsv_ctx_add_gate(ctx, 8, ins[0], ins[1], s)
        lib.gsv_ctx_add_gate(ctx, 0, ins[0], ins[1], c)
        outs[0], outs[1] = s, c

    @BODY
    def root(ctx, user, ins, n_in, outs, arity):
        tmp = (ctypes.c_uint32 * 2)()
        lib.gsv_ctx_component(ctx, b"half_adder", ins, 2, 2, half_adder, None, tmp)
        outs[0] = tmp[0]  # the carry is never read -> its And gate is dead

    h = lib.gsv_program_record(b"ha", 2, 1, root, None, None)
    assert h
    info = gsv._ProgramInfo()
    assert lib.gsv_program_get_info(h, ctypes.byref(info)) == 0
    assert (info.n_gates, info.n_live_gates, info.n_ciphertexts) == (2, 1, 0)
    lib.gsv_program_destroy(h)


def test_no_cpu_fallback(gsv):
    if gsv.device_count() > 0:
        pytest.skip("a GPU is present")
    p = gsv.Program("fq_add")
    with pytest.raises(gsv.GsvError) as e:
        gsv.Session(p, 2)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(gsv.GsvError):
        gsv.commit_labels(np.zeros((4, 16), np.uint8))
    with pytest.raises(gsv.GsvError):
        gsv.hash_blocks(0, np.zeros((4, 16), np.uint8), np.zeros(4, np.uint64))


def test_unknown_circuit(gsv):
    with pytest.raises(gsv.GsvError):
        gsv.Program("no_such_circuit")


def test_host_chain_fold_matches_cbc_mac(gsv):
    """Host half of GSV_CT_COMMIT_HOST vs an independent statement of the chain: h <- AES_K(h ^ ct)
    is CBC-MAC with a zero IV under the fixed key (OpenSSL), for 1..13 interleaved instances, folded
    in two pieces (the drain-buffer boundary)."""
    import numpy as np
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

    rng = np.random.default_rng(3)
    for n_inst in (1, 2, 3, 4, 7, 8, 13):
        blocks = rng.integers(0, 256, (57, n_inst, 16), dtype=np.uint8)
        h = gsv.host_chain_fold(np.zeros((n_inst, 16), np.uint8), blocks[:20])
        h = gsv.host_chain_fold(h, blocks[20:])
        h2 = gsv.host_chain_fold(np.zeros((n_inst, 16), np.uint8), blocks.transpose(1, 0, 2), instance_major=True)
        assert np.array_equal(h, h2)
        # one position at a time (the last round of a step is merged with the next step's whitening inside a
        # call, so single-position calls exercise the unmerged ends), and an empty fold
        h3 = np.zeros((n_inst, 16), np.uint8)
        for p in range(blocks.shape[0]):
            h3 = gsv.host_chain_fold(h3, blocks[p:p + 1])
        assert np.array_equal(h, h3)
        assert np.array_equal(gsv.host_chain_fold(h, blocks[:0]), h)
        for i in range(n_inst):
            enc = Cipher(algorithms.AES(bytes([0x42]) * 16), modes.CBC(bytes(16))).encryptor()
            want = enc.update(blocks[:, i, :].tobytes())[-16:]
            assert bytes(h[i]) == want
    # the drain layout of the device ring: [quad][position][4 chains][16 bytes]
    for n_quads in (1, 2, 3, 4, 7):
        rows = rng.integers(0, 256, (n_quads, 41, 4, 16), dtype=np.uint8)
        h = gsv.host_chain_fold_quads(np.zeros((4 * n_quads, 16), np.uint8), rows[:, :17])
        h = gsv.host_chain_fold_quads(h, rows[:, 17:])
        for q in range(n_quads):
            for j in range(4):
                enc = Cipher(algorithms.AES(bytes([0x42]) * 16), modes.CBC(bytes(16))).encryptor()
                assert bytes(h[4 * q + j]) == enc.update(rows[q, :, j, :].tobytes())[-16:]
