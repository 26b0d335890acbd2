"""Third-party arithmetic pinned with vectors that do NOT come from this repository's own restatements
(tests/golden/third_party_kats.json, tests/golden/seahash_buffer_mode.json; generator and provenance:
tests/golden/gen_third_party_kats.py).

  * SplitMix64 / xoshiro256++      published known-answer vectors (Vigna's reference C code as quoted by rand_xoshiro's
                                   own unit tests)           -> oracle Xoshiro256pp AND the engine's ParticleRng
  * seahash 4.1                    the crate's documented vector + 97 lengths computed by a restatement of the crate's
                                   BUFFER form (4 fixed lanes) -> oracle (stream form), bgr_seahash, and the CUDA kernels
  * Duration::as_secs_f32          IEEE binary32 evaluation of the std formula -> oracle and bgr_ggrs_time_delta_bits
  * rand 0.9 f32 random_range      the retry / shrink-scale branch is proven unreachable for (-200, 200), so the
                                   restated straight-line formula is the whole algorithm for the example's ranges
"""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

from bevy_ggrs_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "third_party_kats.json")))
BUF = json.load(open(os.path.join(HERE, "golden", "seahash_buffer_mode.json")))


def _pattern(n):
    return bytes((((i * 37 + 11) & 0xFF) | (0xC0 if i % 5 == 0 else 0)) & 0xFF for i in range(n))


def test_goldens_are_what_the_committed_generator_writes(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(HERE, "golden", "gen_third_party_kats.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert [v["hash"] for v in BUF["vectors"]] == [hex(gen.seahash_buffer_mode(gen.pattern(n))) for n in range(97)]
    assert gen.pattern(40) == _pattern(40)


def test_splitmix64_published_vectors_product_and_oracle(oracle_lib):
    lib = capi.load_library()
    for v in KATS["splitmix64"]:
        n = len(v["next_u64"])
        out = (C.c_uint64 * n)()
        assert lib.bgr_splitmix64_stream(v["seed"], n, out) == 0
        assert list(out) == v["next_u64"], v["source"]
        # Xoshiro256PlusPlus::seed_from_u64 = four SplitMix64 outputs: oracle and product state words
        st = (C.c_uint64 * 4)()
        oracle_lib.orc_xoshiro_seed_state(v["seed"], st)
        assert list(st) == v["next_u64"][:4]


def test_xoshiro256plusplus_published_vectors_product_and_oracle(oracle_lib):
    lib = capi.load_library()
    for v in KATS["xoshiro256plusplus"]:
        n = len(v["next_u64"])
        st = (C.c_uint64 * 4)(*v["state"])
        out = (C.c_uint64 * n)()
        assert lib.bgr_particle_rng_stream(0, st, n, out, None, 0.0, 1.0) == 0
        assert list(out) == v["next_u64"], v["source"]
        out2 = (C.c_uint64 * n)()
        oracle_lib.orc_xoshiro_from_state(st, n, out2)
        assert list(out2) == v["next_u64"]


def test_seeded_streams_of_product_and_oracle_agree_and_follow_the_f32_formula(oracle_lib):
    """seed_from_u64(123) (particles.rs:199): raw stream and random_range(-200..200) of the engine's ParticleRng equal
    the oracle's, and both equal the straight-line formula evaluated in numpy float32 on the pinned raw stream."""
    lib = capi.load_library()
    n = 256
    u, f = (C.c_uint64 * n)(), (C.c_float * n)()
    assert lib.bgr_particle_rng_stream(123, None, n, u, f, -200.0, 200.0) == 0
    uo, fo = (C.c_uint64 * n)(), (C.c_float * n)()
    oracle_lib.orc_xoshiro_stream(123, n, uo, fo, -200.0, 200.0)
    assert list(u) == list(uo)
    raw = np.array(list(u), dtype=np.uint64)
    bits = ((raw >> np.uint64(32)).astype(np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)
    v01 = bits.view(np.float32) - np.float32(1.0)
    want = (v01 * np.float32(400.0)).astype(np.float32) + np.float32(-200.0)
    assert np.array_equal(np.array(list(f), dtype=np.float32).view(np.uint32), want.astype(np.float32).view(np.uint32))
    assert np.array_equal(np.array(list(fo), dtype=np.float32).view(np.uint32), want.view(np.uint32))


def test_rand_uniform_f32_retry_branch_is_unreachable_for_the_examples_range():
    """UniformFloat::sample_single only loops (and shrinks `scale`) when value0_1 * scale + low >= high.  Over ALL 2^23
    mantissas of value1_2 the result for (-200, 200) stays inside [-200, 200): the straight-line formula is complete."""
    bits = np.arange(1 << 23, dtype=np.uint32) | np.uint32(0x3F800000)
    v01 = bits.view(np.float32) - np.float32(1.0)
    res = (v01 * np.float32(400.0)).astype(np.float32) + np.float32(-200.0)
    assert res.dtype == np.float32
    assert float(res.max()) < 200.0 and float(res.min()) >= -200.0
    assert float(res.max()) == float(np.float32(199.99994))


def test_seahash_documented_vector_everywhere(oracle_lib):
    lib = capi.load_library()
    for v in KATS["seahash"]:
        b = v["ascii"].encode()
        buf = C.create_string_buffer(b, len(b))
        assert lib.bgr_seahash(buf, len(b)) == v["hash"]
        assert oracle_lib.orc_seahash(buf, len(b)) == v["hash"]


def test_seahash_buffer_form_vectors_all_lengths_product_and_oracle(oracle_lib):
    """lengths 0..96: every excess-byte arm of buffer.rs' match (0, 1..=7, 8, 9..=15, 16, 17..=23, 24, 25..=31) with
    0, 1, 2 and 3 full 32-byte blocks in front, against the stream-form restatements."""
    lib = capi.load_library()
    for v in BUF["vectors"]:
        b = _pattern(v["len"])
        buf = C.create_string_buffer(b, len(b))
        assert lib.bgr_seahash(buf, len(b)) == int(v["hash"], 16), v["len"]
        assert oracle_lib.orc_seahash(buf, len(b)) == int(v["hash"], 16), v["len"]


def test_integer_field_writes_equal_the_byte_stream(oracle_lib):
    """`Hash for u32 / u64` appends little-endian bytes (seahash 4.x tail buffer): field-wise hashing of the
    buffer-form vectors' bytes gives the same value — the shape of the particles hashers (3 x u32, particles.rs:107-120)."""
    for n_words in range(1, 9):
        b = _pattern(4 * n_words)
        arr = (C.c_uint32 * n_words)(*struct.unpack("<%dI" % n_words, b))
        assert oracle_lib.orc_seahash_u32_fields(arr, n_words) == int(BUF["vectors"][4 * n_words]["hash"], 16)


def test_duration_as_secs_f32_vectors(oracle_lib):
    lib = capi.load_library()
    for v in KATS["duration_as_secs_f32"]["vectors"]:
        assert lib.bgr_ggrs_time_delta_bits(v["fps"], v["frame"]) == v["bits"], v
        assert oracle_lib.orc_ggrs_time_delta_bits(v["fps"], v["frame"]) == v["bits"], v
    by = {(v["fps"], v["frame"]): v for v in KATS["duration_as_secs_f32"]["vectors"]}
    assert by[(60, 1)]["delta_ns"] == 16_666_666 and by[(60, 1)]["bits"] == 0x3C888888
    assert by[(60, 2)]["delta_ns"] == 16_666_667 and by[(60, 2)]["bits"] == 0x3C888889
    assert by[(1, 1)]["bits"] == 0x3F800000                    # a whole second: secs = 1, nanos = 0


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [capi.BGR_CFG_FORCE_STEPWISE, 0])
def test_gpu_generic_hash_kernels_reproduce_the_buffer_form_vectors(flags):
    """The CUDA byte-range hashers (stepwise: k_image_tma; default: k_generic_program's hash_row_range, word-aligned
    and byte-wise arms) on one entity per length:
    checksum = entity_part(1, 1) ^ seahash(seahash(order 0 ‖ custom)) with custom = the buffer-form vector."""
    from bevy_ggrs_b200.engine import Engine
    from bevy_ggrs_b200.session import SAVE, Request
    lib = capi.load_library()

    def h(b):
        buf = C.create_string_buffer(b, len(b))
        return lib.bgr_seahash(buf, len(b))

    for length in list(range(1, 41)) + [47, 48, 63, 64, 65, 96]:
        custom = int(BUF["vectors"][length]["hash"], 16)
        eng = Engine(max_entities=4, max_depth=2, flags=flags)
        col = eng.rollback_component("Blob", length)
        eng.checksum_component(col, 0, length)
        eng.build()
        eng.spawn(1)
        eng.write_component(col, 0, np.frombuffer(_pattern(length), dtype=np.uint8).reshape(1, length))
        (frame, cs), = eng.handle_requests((capi.BGR_SESSION_NONE, 0, 0, 0), [Request(SAVE, 0)])
        per_entity = h(struct.pack("<QQ", 0, custom))
        want = h(struct.pack("<QQ", 1, 1)) ^ h(struct.pack("<Q", per_entity))
        assert cs == want, length
        eng.close()


@pytest.mark.gpu
def test_gpu_fused_particles_hash_reproduces_the_buffer_form_composition():
    """The fused kernel's specialised 12-byte hash (z == 0 fast path and general path) for single entities whose
    translation / velocity bytes are the KAT pattern: expected value composed from the buffer-form restatement."""
    import importlib.util
    from bevy_ggrs_b200.engine import Engine
    from bevy_ggrs_b200.session import SAVE, Request
    from bevy_ggrs_b200.stress import register_particles
    spec = importlib.util.spec_from_file_location("gen", os.path.join(HERE, "golden", "gen_third_party_kats.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    sh = gen.seahash_buffer_mode
    for z in (0.0, 3.5):
        n = 5
        tf = np.zeros((n, 10), np.float32)
        vel = np.zeros((n, 3), np.float32)
        rng = np.random.default_rng(2024)
        tf[:, :2] = rng.uniform(-300, 300, (n, 2)); vel[:, :2] = rng.uniform(-100, 100, (n, 2))
        tf[:, 2] = z; vel[:, 2] = -z
        tf[:, 6] = 1; tf[:, 7:] = 1
        eng = Engine(max_entities=n, max_depth=2)
        t, v, l = register_particles(eng)
        eng.build()
        eng.spawn(n)
        eng.write_component(t, 0, tf); eng.write_component(v, 0, vel)
        eng.write_component(l, 0, np.full(n, 99, np.uint64))
        (frame, cs), = eng.handle_requests((capi.BGR_SESSION_NONE, 0, 0, 0), [Request(SAVE, 0)])
        assert eng.last_path_fused()
        xt = xv = 0
        for i in range(n):
            xt ^= sh(struct.pack("<QQ", i, sh(tf[i, :3].tobytes())))
            xv ^= sh(struct.pack("<QQ", i, sh(vel[i].tobytes())))
        want = sh(struct.pack("<QQ", n, n)) ^ sh(struct.pack("<Q", xt)) ^ sh(struct.pack("<Q", xv))
        assert cs == want, z
        eng.close()
