"""CPU emulation of the CUDA tile algorithm (tests/emul/emul.cpp, which drives the
very device functions of csrc/hb_core.cuh) checked against the oracle.  This is
how the algorithm is validated in the GPU-less build container; the -m gpu
tests repeat the same cases on the real kernels."""
import numpy as np
import pytest

import emul_lib as E
import oracle_lib as O
import huffmandecoderongpus_b200 as hb

SHAPES_BIG = [(4, 256), (8, 256), (16, 256)]
SHAPES_SMALL = [(1, 4), (2, 8), (4, 32), (1, 64)]


def _stream(name):
    p = O.corpus_path(name)
    if p is None:
        pytest.skip(f"{name} corpus not present")
    return O.load_huff(p)


@pytest.mark.parametrize("name", list(O.CORPORA))
@pytest.mark.parametrize("shape", SHAPES_BIG)
def test_corpora(name, shape):
    st = _stream(name)
    got, stats, rc = E.decode(st, *shape)
    assert rc == 0
    assert got.size == st.usize
    assert O.sha256(got) == O.CORPORA[name][2]


@pytest.mark.parametrize("name", ["hello", "paper1", "news"])
@pytest.mark.parametrize("shape", SHAPES_SMALL)
def test_small_tiles_many_boundaries(name, shape):
    st = _stream(name)
    got, stats, rc = E.decode(st, *shape)
    assert rc == 0 and np.array_equal(got, O.simple_decode(st))
    if name != "hello":
        assert stats["tiles_entry_nonzero"] > 0  # the entry fix-up path really ran


@pytest.mark.parametrize("shape", [(4, 256), (1, 4)])
def test_prefix_sweep(shape):
    # reference graphtest / setTargetSizes, framework/mainrun.c:361-410
    st = _stream("paper1")
    lut = hb.build_lut(st.tree)
    for target in list(range(1, 70)) + list(range(10000, 266692, 31337)) + [st.bits - 1, st.bits]:
        bits, usize = O.prefix_sizes(st, target)
        w = E.words_of(st.data, (bits + 7) // 8)
        out, _, res, _, rc = E.run(lut, w, bits, bits, *shape)
        assert rc == 0 and int(res[0]) == usize, target
        assert np.array_equal(out[:usize], O.simple_decode(st, bits=bits)), target


@pytest.mark.parametrize("shape", [(4, 256), (2, 8)])
def test_stream_cut_inside_a_codeword(shape):
    # the serial oracle emits a symbol only on reaching a leaf: a trailing partial
    # codeword yields nothing (framework/mainrun.c:44-53)
    st = _stream("paper1")
    lut = hb.build_lut(st.tree)
    for bits in [1, 2, 3, 5, 31, 32, 33, 127, 128, 129, 1000, 4095, 4096, 4097, 32767, 32768, 32769,
                 100001, st.bits - 3]:
        want = O.simple_decode(st, bits=bits)
        w = E.words_of(st.data, (bits + 7) // 8)
        # bits past the cut are garbage (the rest of the real stream), not zeros
        w2 = E.words_of(st.data, min(st.nbytes, (bits + 7) // 8 + 8))
        for words in (w, w2):
            out, _, res, _, rc = E.run(lut, words, bits, bits, *shape)
            assert rc == 0 and int(res[0]) == want.size, bits
            assert np.array_equal(out[: want.size], want), bits


def test_empty_stream():
    st = _stream("hello")
    lut = hb.build_lut(st.tree)
    out, smap, res, _, rc = E.run(lut, np.zeros(4, np.uint32), 0, 0)
    assert rc == 0 and int(res[0]) == 0
    assert [int(m) for m in smap] == list(range(32))  # identity map


def _compose(maps, rank):
    cur, base = 0, 0
    for r in range(rank):
        m = int(maps[r][cur])
        base += m >> 8
        cur = m & 31
    return cur, base


@pytest.mark.parametrize("name,cuts", [
    ("paper1", [16, 4096, 4112, 20000]),
    ("news", [8192 * 3, 8192 * 17]),
    ("kjv", [1024 * 1024, 2 * 1024 * 1024 + 16]),
    ("ecoli", [500000 - 500000 % 16]),
])
@pytest.mark.parametrize("shape", [(4, 256), (2, 8)])
def test_byte_range_shards(name, cuts, shape):
    """Multi-GPU decomposition: contiguous 16-byte-aligned byte ranges, one map
    per shard, host-side composition (what hb_shard_compose does), independent
    emit of every shard into its own buffer."""
    if shape == (2, 8) and name in ("kjv", "ecoli"):
        pytest.skip("small shape only on small corpora")
    st = _stream(name)
    want = O.simple_decode(st)
    lut = hb.build_lut(st.tree)
    bounds = [0] + cuts + [st.nbytes]
    shards = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        last = b == st.nbytes
        bits_own = (st.bits - 8 * a) if last else 8 * (b - a)
        halo_end = min(st.nbytes, b + 8)
        bits_avail = bits_own if last else min(st.bits - 8 * a, 8 * (halo_end - a))
        words = E.words_of(st.data[a:], halo_end - a)
        shards.append((words, bits_own, bits_avail))
    maps = []
    for words, bo, ba in shards:
        _, smap, _, _, rc = E.run(lut, words, bo, ba, *shape, emit=False)
        assert rc == 0
        maps.append(smap)
    pieces = []
    for r, (words, bo, ba) in enumerate(shards):
        entry, base = _compose(maps, r)
        out, _, res, _, rc = E.run(lut, words, bo, ba, *shape, entry=entry, base=0, out_offset=r % 5)
        assert rc == 0
        assert int(res[2]) == entry
        pieces.append(out[: int(res[0])])
        assert base == sum(p.size for p in pieces[:-1])
    got = np.concatenate(pieces)
    assert got.size == want.size and np.array_equal(got, want)


@pytest.mark.parametrize("kind,n", [(hb.MODEL_ENGLISH, 300000), (hb.MODEL_FIBONACCI, 300000),
                                    (hb.MODEL_DNA, 100000), (hb.MODEL_UNIFORM8, 20000)])
@pytest.mark.parametrize("shape", [(4, 256), (8, 256), (1, 4)])
def test_synthetic_models(kind, n, shape):
    m = hb.Model(kind)
    if kind == hb.MODEL_FIBONACCI:
        assert 20 < m.maxlen <= 32  # the adversarial tree needs a multi-level table
    f, syms = m.huff_file_cpu(seed=0x48554646, n=n)
    lut = hb.build_lut(f.tree)
    assert lut["maxlen"] == m.maxlen and lut["minlen"] == m.minlen
    got, stats, rc = E.decode(f, *shape, lut=lut)
    assert rc == 0 and np.array_equal(got, syms)
    # and the oracle agrees with the generator
    st = O.Stream(f.tree, f.data, f.bits, f.usize)
    assert np.array_equal(O.simple_decode(st), syms)


def test_output_too_small_is_reported():
    st = _stream("paper1")
    lut = hb.build_lut(st.tree)
    w = E.words_of(st.data, st.nbytes)
    _, _, _, _, rc = E.run(lut, w, st.bits, st.bits, out_capacity=st.usize - 1)
    assert rc == -6


def test_output_alignment_offsets():
    st = _stream("paper1")
    want = O.simple_decode(st)
    lut = hb.build_lut(st.tree)
    w = E.words_of(st.data, st.nbytes)
    for off in range(0, 16):
        out, _, res, _, rc = E.run(lut, w, st.bits, st.bits, out_offset=off)
        assert rc == 0 and np.array_equal(out[: want.size], want)
        assert not out[want.size:].any()  # nothing written past the end


@pytest.mark.parametrize("shape,win", [((4, 256), 4096), ((4, 256), 1008), ((8, 256), 2048), ((2, 8), 16), ((1, 4), 16)])
def test_emit_in_several_windows(shape, win):
    """A staging buffer smaller than a tile's output (data more compressible than
    the code table suggests): the emit kernel's window loop must stitch the tile
    back together exactly."""
    for name in ("paper1", "ecoli"):
        st = _stream(name)
        if name == "ecoli" and shape[1] < 256:
            continue
        lut = hb.build_lut(st.tree)
        w = E.words_of(st.data, st.nbytes)
        for off in (0, 5):
            out, _, res, _, rc = E.run(lut, w, st.bits, st.bits, *shape, emit_win=win, out_offset=off)
            assert rc == 0 and int(res[0]) == st.usize
            assert O.sha256(out[: st.usize]) == O.CORPORA[name][2]
            assert not out[st.usize:].any()


def _bit_walk(tree, node, bits):
    """reference semantics (framework/mainrun.c:38-55) from an arbitrary node: returns
    (node after the bits, symbols completed)"""
    ends = 0
    for b in bits:
        node = int(tree[node]["ione"] if b else tree[node]["izero"])
        if tree[node]["izero"] == -1:
            ends += 1
            node = 0
    return node, ends


@pytest.mark.parametrize("name", ["hello", "paper1", "world192", "ecoli", "fib"])
def test_transducer_table(name):
    """hb_lut.c build_fsm against a bit-serial walk of the node array: states are the
    internal nodes (root = 0), an entry is (state after 8 bits, codewords ended)."""
    tree = hb.Model(hb.MODEL_FIBONACCI).huff_file_cpu(seed=1, n=10)[0].tree if name == "fib" else _stream(name).tree
    lut = hb.build_lut(tree)
    ns = lut["fsm_states"]
    internal = [i for i in range(tree.shape[0]) if tree[i]["izero"] != -1]
    assert ns == len(internal) <= 256
    # recover the node of every state by walking the 1-bit table from the root
    node_of = {0: 0}
    todo = [0]
    while todo:
        s = todo.pop()
        for bit in (0, 1):
            r = int(lut["fsm_bstep"][2 * s + bit])
            child = int(tree[node_of[s]]["ione"] if bit else tree[node_of[s]]["izero"])
            leaf = tree[child]["izero"] == -1
            assert bool(r >> 8) == bool(leaf)
            if not leaf and (r & 0xFF) not in node_of:
                node_of[r & 0xFF] = child
                todo.append(r & 0xFF)
    assert len(node_of) == ns
    state_of = {v: k for k, v in node_of.items()}
    rng = np.random.default_rng(7)
    for s in range(ns):
        d = 0
        v = node_of[s]
        # depth = distance from the root
        depth = {0: 0}
        stack = [0]
        while stack and v not in depth:
            u = stack.pop()
            for c in (int(tree[u]["izero"]), int(tree[u]["ione"])):
                if c != -1 and c not in depth:
                    depth[c] = depth[u] + 1
                    stack.append(c)
        assert lut["fsm_depth"][s] == depth[v]
        for b in (rng.integers(0, 256, 24).tolist() + [0, 255]):
            node, ends = _bit_walk(tree, v, [(b >> i) & 1 for i in range(8)])
            ent = int(lut["fsm"][s * 256 + b])
            assert ent >> 8 == state_of[node] and (ent & 0xFF) == ends, (s, b)
    for r in range(1, 8):   # partial steps from the root
        for x in range(1 << r):
            node, ends = _bit_walk(tree, 0, [(x >> i) & 1 for i in range(r)])
            ent = int(lut["fsm_pstep"][(1 << r) + x])
            assert ent >> 8 == state_of[node] and (ent & 0xFF) == ends, (r, x)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", [(8, 256), (2, 8)])
def test_sync_paths(mode, shape):
    """probe kernel only / transducer kernel on full tiles (product dispatch) / both
    with their records compared tile by tile"""
    st = _stream("news")
    got, stats, rc = E.decode(st, *shape, sync_mode=mode)
    assert rc == 0 and O.sha256(got) == O.CORPORA["news"][2]


# codes whose chains merge late or never, and codes with very long codewords
HARD_CODES = [
    [2, 2, 2, 4, 4, 4, 4],                       # all lengths even: odd offsets never merge
    [7] * 64 + [8] * 128,                        # nearly fixed length
    [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 12],  # minimum length 1
    [3] * 8,                                     # exactly fixed length (closed-form path)
    list(range(1, 31)) + [30],                   # 30-bit codewords: three table levels
    [2] + [3] * 5 + list(range(4, 33)) + [32],   # 32-bit codewords (HB_MAX_CODELEN)
]


@pytest.mark.parametrize("lengths", HARD_CODES)
@pytest.mark.parametrize("shape", [(4, 256), (8, 256), (1, 4)])
def test_badly_synchronising_codes(lengths, shape):
    """many stitch rounds, unmerged hypotheses, marker entries and multi-level probes"""
    tree, codes = O.tree_from_lengths(lengths)
    rng = np.random.default_rng(len(lengths))
    # long codewords must actually occur: draw symbols uniformly, not by code probability
    syms = rng.integers(0, len(lengths), 40000 if shape[1] == 256 else 3000).astype(np.uint8)
    data, bits = O.encode_with_codes(codes, syms)
    st = O.Stream(tree, data, bits, syms.size)
    assert np.array_equal(O.simple_decode(st), syms)
    got, stats, rc = E.decode(st, *shape)
    assert rc == 0 and np.array_equal(got, syms)


@pytest.mark.parametrize("mode", [0, 1, 3, 4, 5])
@pytest.mark.parametrize("shape", [(4, 256), (8, 256), (16, 256), (2, 8)])
@pytest.mark.parametrize("name", ["book2", "world192"])
def test_emit_paths(name, mode, shape):
    """byte-store emit walk (E-table) and the word-store walks (E64-table, E32-table) give the
    same bytes, at every output alignment"""
    st = _stream(name)
    lut = hb.build_lut(st.tree)
    w = E.words_of(st.data, st.nbytes)
    for off in (0, 1, 2, 3):
        out, _, res, _, rc = E.run(lut, w, st.bits, st.bits, *shape, emit_mode=mode, out_offset=off, ep_wf=0)
        assert rc == 0 and int(res[0]) == st.usize
        assert O.sha256(out[: st.usize]) == O.CORPORA[name][2]
        assert not out[st.usize:].any()


def test_e64_table():
    """E64 entries against the E-/S-table semantics: up to four symbols, the window
    selector, 8 * nsym, bits consumed, marker where no codeword fits"""
    marks = 0
    for tree in (_stream("world192").tree,                       # 20-bit codes: markers exist, 12-bit index
                 hb.Model(hb.MODEL_FIBONACCI).tree):             # short codes: 11-bit index
        marks += _check_e64(tree)
    assert marks > 0


def _check_e64(tree):
    class _S:   # the loop below only needs .tree
        pass
    st = _S()
    st.tree = tree
    lut = hb.build_lut(st.tree)
    wf64 = lut["wf64"]
    assert wf64 == (11 if lut["implied_avg_len"] <= 3.5 else 12)
    e64 = lut["e64"].reshape(-1, 2)
    assert e64.shape[0] == 1 << wf64
    marks = 0
    for x in range(1 << wf64):
        syms, meta = int(e64[x, 0]), int(e64[x, 1])
        ns, bits = (meta >> 19) & 7, meta >> 26
        assert (meta >> 16) & 0x3FF == 8 * ns and (meta & 0xFFFF) == 0x3210 + 0x1111 * ns
        # greedy walk over the wf64 index bits: how many whole codewords fit (at most four)
        node, pos, fit, last = 0, 0, 0, 0
        while pos < wf64 and fit < 4:
            node = int(st.tree[node]["ione"] if (x >> pos) & 1 else st.tree[node]["izero"])
            pos += 1
            if st.tree[node]["izero"] == -1:
                fit, last, node = fit + 1, pos, 0
        if fit == 0:
            assert ns == 0 and bits == 48 and syms == 0
            marks += 1
            continue
        assert ns == fit and bits == last
        # walk the tree over the index bits
        node, pos, got = 0, 0, []
        while len(got) < ns:
            node = int(st.tree[node]["ione"] if (x >> pos) & 1 else st.tree[node]["izero"])
            pos += 1
            if st.tree[node]["izero"] == -1:
                got.append(int(st.tree[node]["sym"]))
                node = 0
        assert pos == bits and got == [(syms >> (8 * i)) & 0xFF for i in range(ns)]
        assert syms >> (8 * ns) == 0 if ns < 4 else True
    return marks


@pytest.mark.parametrize("shape", [(8, 256), (2, 8)])
def test_tree_with_more_than_256_internal_nodes(shape):
    """300 leaves (symbols repeat): no transducer table, the probe sync kernel does all tiles"""
    lengths = [8] * 212 + [9] * 88
    tree, codes = O.tree_from_lengths(lengths)
    lut = hb.build_lut(tree)
    assert lut["fsm_states"] == 0 and tree.shape[0] == 599
    rng = np.random.default_rng(3)
    syms = rng.integers(0, 300, 30000)
    data, bits = O.encode_with_codes(codes, syms)
    st = O.Stream(tree, data, bits, syms.size)
    want = (syms & 255).astype(np.uint8)
    assert np.array_equal(O.simple_decode(st), want)
    got, stats, rc = E.decode(st, *shape, lut=lut)
    assert rc == 0 and np.array_equal(got, want)


@pytest.mark.parametrize("seed", range(4))
def test_random_codes_and_streams(seed):
    """the CPU twin of tests/test_gpu_stress.py: random complete trees (codewords of up to 32
    bits), random streams, every tile shape and both emit walks, against the oracle"""
    rng = np.random.default_rng(2000 + seed)
    for case in range(12):
        nleaves = int(rng.choice([2, 3, 5, 17, 64, 200, 256]))
        maxlen = int(rng.choice([4, 9, 13, 20, 32]))
        lengths = O.random_lengths(rng, nleaves, maxlen)
        tree, codes = O.tree_from_lengths(lengths)
        w = np.array([2.0 ** (-l) for l in lengths])
        mode = int(rng.integers(3))
        p = np.ones(len(lengths)) if mode == 0 else (w if mode == 1 else 1.0 / w)
        n = int(rng.choice([1, 7, 300, 5000, 40000]))
        syms = rng.choice(len(lengths), size=n, p=p / p.sum())
        data, bits = O.encode_with_codes(codes, syms)
        st = O.Stream(tree, data, bits, n)
        want = (syms & 255).astype(np.uint8)
        shape = [(4, 256), (8, 256), (16, 256), (2, 8), (1, 4)][int(rng.integers(5))]
        lut = hb.build_lut(tree)
        wds = E.words_of(st.data, (bits + 7) // 8)
        out, _, res, _, rc = E.run(lut, wds, bits, bits, *shape, emit_mode=int(rng.choice([0, 1, 3, 4, 5, 5])),
                                   ep_wf=int(rng.choice([0, 9, 10, 12, 13, 14])), out_offset=int(rng.integers(16)))
        tag = (seed, case, nleaves, maxlen, n, shape)
        assert rc == 0 and int(res[0]) == n, tag
        assert np.array_equal(out[:n], want), tag
        assert not out[n:].any(), tag


# ---- word-store walk with 32-bit table entries (hb_emit_words32 / hb_emit32_kernel) ------

@pytest.mark.parametrize("ep_wf", [0, 9, 11, 14])
@pytest.mark.parametrize("shape", [(8, 256), (4, 32), (16, 256), (2, 8)])
@pytest.mark.parametrize("name", ["paper1", "world192", "ecoli", "kjv"])
def test_e32_emit_corpora(name, shape, ep_wf):
    """E32-table probes (three symbols, funnel-shift window update), unclipped / clipped phases
    of the last word, markers (index narrower than the longest codeword), every alignment"""
    st = _stream(name)
    if name == "kjv" and (shape != (8, 256) or ep_wf != 0):
        pytest.skip("large corpus: product shape only")
    lut = hb.build_lut(st.tree)
    w = E.words_of(st.data, st.nbytes)
    for off in ((0, 1, 2, 3, 7, 13) if name == "paper1" else (0, 5)):
        out, _, res, _, rc = E.run(lut, w, st.bits, st.bits, *shape, emit_mode=3, ep_wf=ep_wf, out_offset=off)
        assert rc == 0 and int(res[0]) == st.usize
        assert O.sha256(out[: st.usize]) == O.CORPORA[name][2]
        assert not out[st.usize:].any()


def test_e32_emit_truncated_and_windows():
    """stream tails (partial subsequences: hb_emit_clipped32), cut-off last codewords and tiles
    emitted in several staging windows"""
    st = _stream("paper1")
    lut = hb.build_lut(st.tree)
    w = E.words_of(st.data, st.nbytes)
    full = O.simple_decode(st)
    for bits in (st.bits - 1, st.bits - 7, st.bits // 2 + 3, 8 * 4096 + 5, 777):
        sub = O.Stream(st.tree, st.data, bits, 0)
        want = O.simple_decode(sub)
        out, _, res, _, rc = E.run(lut, w, bits, bits, 8, 256, emit_mode=3, ep_wf=0)
        assert rc == 0 and int(res[0]) == want.size
        assert np.array_equal(out[: want.size], want) and np.array_equal(want, full[: want.size])
    for win in (2048, 1008, 144):
        for mode in (3, 4, 5):  # modes 4, 5: warp tiles (hb_emit32w_kernel, one / two subsequences per lane)
            out, _, res, _, rc = E.run(lut, w, st.bits, st.bits, 8, 256, emit_win=win, out_offset=3, emit_mode=mode, ep_wf=0)
            assert rc == 0 and int(res[0]) == st.usize and np.array_equal(out[: st.usize], full)
    for bits in (st.bits - 7, 8 * 4096 + 5, 777):
        want = O.simple_decode(O.Stream(st.tree, st.data, bits, 0))
        for mode in (4, 5):
            out, _, res, _, rc = E.run(lut, w, bits, bits, 8, 256, emit_mode=mode, ep_wf=0, out_offset=5)
            assert rc == 0 and int(res[0]) == want.size and np.array_equal(out[: want.size], want)


# ---- flat emit walk (hb_emit_flat / hb_emitf_kernel) -----------------------------------

@pytest.mark.parametrize("ep_wf", [8, 10, 11, 12])
@pytest.mark.parametrize("shape", [(8, 256), (4, 32), (16, 256)])
@pytest.mark.parametrize("name", ["paper1", "book2", "world192", "ecoli", "kjv"])
def test_flat_emit_corpora(name, shape, ep_wf):
    """flat walk over every tile but the last: whole staging words only, final word repaired
    after the barrier; every output alignment"""
    st = _stream(name)
    if name in ("kjv", "book2") and (shape != (8, 256) or ep_wf != 10):
        pytest.skip("large corpus: product shape only")
    lut = hb.build_lut(st.tree)
    w = E.words_of(st.data, st.nbytes)
    for off in ((0, 1, 2, 3, 7, 13) if name == "paper1" else (0, 5)):
        out, _, res, _, rc = E.run(lut, w, st.bits, st.bits, *shape, emit_mode=2, ep_wf=ep_wf, out_offset=off)
        assert rc == 0 and int(res[0]) == st.usize
        assert O.sha256(out[: st.usize]) == O.CORPORA[name][2]
        assert not out[st.usize:].any()


@pytest.mark.parametrize("win", [2048, 1008, 144])
def test_flat_emit_in_several_windows(win):
    for name in ("paper1", "ecoli"):
        st = _stream(name)
        lut = hb.build_lut(st.tree)
        w = E.words_of(st.data, st.nbytes)
        for off in (0, 5):
            out, _, res, _, rc = E.run(lut, w, st.bits, st.bits, 8, 256, emit_win=win, out_offset=off, emit_mode=2)
            assert rc == 0 and int(res[0]) == st.usize
            assert O.sha256(out[: st.usize]) == O.CORPORA[name][2]
            assert not out[st.usize:].any()


@pytest.mark.parametrize("seed", range(6))
def test_flat_emit_random_codes(seed):
    """random complete trees with codewords of up to 32 bits: the look-ahead rows and the
    single-symbol fallback of the flat walk"""
    rng = np.random.default_rng(7000 + seed)
    for case in range(10):
        nleaves = int(rng.choice([2, 3, 5, 17, 64, 200, 256]))
        maxlen = int(rng.choice([4, 9, 13, 20, 32]))
        lengths = O.random_lengths(rng, nleaves, maxlen)
        tree, codes = O.tree_from_lengths(lengths)
        if min(lengths) == max(lengths):
            continue
        wts = np.array([2.0 ** (-l) for l in lengths])
        mode = int(rng.integers(3))
        p = np.ones(len(lengths)) if mode == 0 else (wts if mode == 1 else 1.0 / wts)
        n = int(rng.choice([3000, 20000, 60000]))
        syms = rng.choice(len(lengths), size=n, p=p / p.sum())
        data, bits = O.encode_with_codes(codes, syms)
        st = O.Stream(tree, data, bits, n)
        want = (syms & 255).astype(np.uint8)
        shape = [(8, 256), (4, 32), (8, 256)][int(rng.integers(3))]
        lut = hb.build_lut(tree)
        wds = E.words_of(st.data, (bits + 7) // 8)
        out, _, res, _, rc = E.run(lut, wds, bits, bits, *shape, emit_mode=2, ep_wf=int(rng.choice([8, 10, 11])),
                                   out_offset=int(rng.integers(16)))
        tag = (seed, case, nleaves, maxlen, n, shape)
        assert rc == 0 and int(res[0]) == n, tag
        assert np.array_equal(out[:n], want), tag
        assert not out[n:].any(), tag


# ---- codes whose lengths share a factor: entry offsets off the residue class never occur ----

@pytest.mark.parametrize("lengths", [
    [2, 2, 2, 4, 4, 4, 4],                       # all even (the documented 18x cliff of round 1)
    [2] * 3 + [4] * 3 + [6] * 3 + [8] * 3 + [10] * 3 + [12] * 3 + [14] * 3 + [16] * 4,   # even, long tail
    [4] * 15 + [8] * 16,                         # multiples of 4
    [8] * 255 + [16] * 256,                      # multiples of 8, more than 256 internal nodes
    [3, 3, 3, 3, 3, 3, 3, 6, 6, 6, 6, 6, 6, 6, 6],   # multiples of 3: no power of two, full path
    [6] * 63 + [12] * 64,                        # gcd 6 -> only the factor 2 is used
])
@pytest.mark.parametrize("shape", [(8, 256), (4, 32), (1, 4)])
def test_codes_with_a_common_length_factor(lengths, shape):
    tree, codes = O.tree_from_lengths(lengths)
    lut = hb.build_lut(tree)
    g = 0
    for l in lengths:
        g = np.gcd(g, l)
    assert lut["len_gcd"] == g
    rng = np.random.default_rng(sum(lengths))
    n = 60000 if shape[1] == 256 else 4000
    syms = rng.integers(0, len(lengths), n)
    data, bits = O.encode_with_codes(codes, syms)
    st = O.Stream(tree, data, bits, n)
    want = (syms & 255).astype(np.uint8)
    for mode in (1, 2, 3, 4, 5):
        got, stats, rc = E.decode(st, *shape, lut=lut, emit_mode=mode)
        assert rc == 0 and np.array_equal(got, want), (mode,)
    # byte-range shards of the same stream (shards start at multiples of 128 bits)
    w = E.words_of(st.data, st.nbytes)
    nb = (bits + 7) // 8
    cut = (nb // 2) // 16 * 16
    if cut >= 16 and shape[1] == 256:
        outs = []
        entry, base = 0, 0
        for (a, b) in ((0, cut), (cut, nb)):
            last = b == nb
            own = bits - 8 * a if last else 8 * (b - a)
            avail = own if last else min(bits - 8 * a, 8 * (b + 16 - a))
            for origin in (a, None):    # the shard's position in the stream known / unknown
                out, smap, res, _, rc = E.run(lut, w[a // 4:], own, avail, *shape, entry=entry, base=base,
                                              origin_byte=origin)
                assert rc == 0
                assert np.array_equal(out[: int(res[0])], want[base: base + int(res[0])])
            outs.append(out[: int(res[0])].copy())
            m = int(smap[entry])
            entry, base = m & 31, base + (m >> 8)
            if not last:
                assert (8 * b + entry) % g == 0     # codeword starts stay in one residue class
        assert np.array_equal(np.concatenate(outs), want)
