"""Pins the oracle's SeaHash restatement (oracle/seahash.hpp).

External pins: the seahash crate's documented vectors.  Cross-restatement pins: the derived
vectors of SURVEY.md §8c (computed by an independent restatement), stored in
tests/golden/seahash_vectors.json.  The reference itself holds no numeric checksum golden.
"""
import ctypes as C
import json
import os
import struct

import numpy as np

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "seahash_vectors.json")))


def _h(lib, b: bytes) -> int:
    buf = C.create_string_buffer(b, len(b))
    return lib.orc_seahash(buf, len(b))


def test_crate_documented_vectors(oracle_lib):
    # seahash crate docs / tests
    assert _h(oracle_lib, b"to be or not to be") == 1988685042348123509
    assert _h(oracle_lib, b"") == 14492805990617963705


def test_survey_derived_vectors(oracle_lib):
    for v in GOLD["bytes_vectors"]:
        assert _h(oracle_lib, bytes.fromhex(v["hex"])) == int(v["hash"], 16), v["name"]


def test_checksum_part_from_value_u32(oracle_lib):
    # ChecksumPart::from_value(&42u32), checksum.rs:38-44 ; self-equality is the reference's own test (:108-113)
    a = oracle_lib.orc_checksum_part_from_u32(42)
    assert a == oracle_lib.orc_checksum_part_from_u32(42)
    assert a == 0x352173BD5A4BA44B


def test_stream_equals_buffer(oracle_lib):
    """Integer writes append little-endian bytes to one stream (seahash 4.x tail buffering):
    hashing 3 x u32 field-by-field equals hashing the 12 concatenated bytes."""
    rng = np.random.default_rng(7)
    for n in range(0, 12):
        vals = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        arr = (C.c_uint32 * max(1, n))(*[int(x) for x in vals])
        assert oracle_lib.orc_seahash_u32_fields(arr, n) == _h(oracle_lib, vals.tobytes())
    for n in range(0, 9):
        vals = rng.integers(0, 2**63, size=n, dtype=np.uint64)
        arr = (C.c_uint64 * max(1, n))(*[int(x) for x in vals])
        assert oracle_lib.orc_seahash_u64_fields(arr, n) == _h(oracle_lib, vals.tobytes())


def test_all_tail_lengths_against_python_restatement(oracle_lib):
    """Every tail length 0..40 against a second, pure-Python restatement of the published algorithm."""
    M = (1 << 64) - 1
    P = 0x6EED0E9DA4D94A4F

    def diffuse(x):
        x = (x * P) & M
        x ^= (x >> 32) >> (x >> 60)
        return (x * P) & M

    def ref(b):
        s = [0x16F11FE89B0D677C, 0xB480A793D8E6C86C, 0x6FE2E5AAF078EBC9, 0x14F994A4C5259381]
        i = 0
        while len(b) - i >= 8:
            t = diffuse(s[0] ^ int.from_bytes(b[i:i + 8], "little"))
            s = [s[1], s[2], s[3], t]
            i += 8
        a = s[0]
        if len(b) - i:
            a = diffuse(a ^ int.from_bytes(b[i:], "little"))
        return diffuse(a ^ s[1] ^ s[2] ^ s[3] ^ len(b))

    rng = np.random.default_rng(11)
    for n in range(0, 41):
        b = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert _h(oracle_lib, b) == ref(b), n


def test_entity_and_component_part_goldens(oracle_lib):
    g = GOLD["composed"]
    # entity part (active, total), entity_checksum.rs:35-43
    assert _h(oracle_lib, struct.pack("<QQ", 2, 2)) == int(g["entity_part_2_2"], 16)
    assert _h(oracle_lib, struct.pack("<QQ", 10**6, 10**6)) == int(g["entity_part_1e6_1e6"], 16)
    # translation (1,2,3): custom hash -> per-entity (order 0) -> part
    custom = _h(oracle_lib, struct.pack("<fff", 1.0, 2.0, 3.0))
    assert custom == int(g["translation_123_custom"], 16)
    per_entity = _h(oracle_lib, struct.pack("<QQ", 0, custom))
    assert per_entity == int(g["translation_123_entity_order0"], 16)
    assert _h(oracle_lib, struct.pack("<Q", per_entity)) == int(g["translation_123_part"], 16)
    assert _h(oracle_lib, struct.pack("<Q", 0)) == int(g["component_part_zero_entities"], 16)
    assert _h(oracle_lib, struct.pack("<I", 0)) == int(g["framecount_0_part"], 16)


def test_ggrs_time_dt_sequence(oracle_lib):
    """time.rs:63-76: at 60 fps the step ending at frame k has 16_666_666 ns if (k-1)%3==0 else ..667
    => delta_secs bits 0x3c888888 / 0x3c888889."""
    bits = [oracle_lib.orc_ggrs_time_delta_bits(60, k) for k in range(1, 13)]
    assert bits == GOLD["dt_bits_60fps_frames_1_to_12"]
    assert set(bits) == {0x3C888888, 0x3C888889}
    for k in range(1, 200):
        ns = (k * 10**9) // 60 - ((k - 1) * 10**9) // 60
        want = np.float32(np.float32(ns) / np.float32(1e9)).view(np.uint32)
        assert oracle_lib.orc_ggrs_time_delta_bits(60, k) == int(want)
