"""One process, N devices (hb_multi_*, b200ApproachMulti): byte-range shards, maps exchanged
by peer copies.  Runs with however many GPUs the box has (1 works: a single shard); on a
multi-GPU box the 2-, 4- and 8-device splits are exercised as well."""
import numpy as np
import pytest

import oracle_lib as O
import huffmandecoderongpus_b200 as hb

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SEED = 0x48554646


def _counts():
    n = torch.cuda.device_count()
    return [k for k in (1, 2, 3, 4, 8) if k <= n]


def _stream(name):
    p = O.corpus_path(name)
    if p is None:
        pytest.skip(f"{name} corpus not present")
    return hb.HuffFile.load(p)


@pytest.mark.parametrize("name", ["hello", "paper1", "world192", "kjv", "ecoli", "bible"])
def test_multi_host_buffers(name):
    """hb_multi_decode_host: upload, shard maps, peer exchange, compose, emit, download"""
    f = _stream(name)
    for n in _counts():
        m = hb.Multi(n)
        out = np.zeros(f.usize + 16, dtype=np.uint8)
        res = m.decode_host(f.tree, f.data, f.bits, out[: f.usize])
        assert res["n_symbols"] == f.usize, (n, res)
        assert O.sha256(out[: f.usize]) == O.CORPORA[name][2], n
        assert not out[f.usize:].any()
        assert sum(res["shard_symbols"]) == f.usize
        # a second call on the same object (cached code tables and buffers)
        out[:] = 0
        res = m.decode_host(f.tree, f.data, f.bits, out[: f.usize])
        assert O.sha256(out[: f.usize]) == O.CORPORA[name][2], n
        m.close()


def test_multi_output_too_small():
    f = _stream("paper1")
    m = hb.Multi(_counts()[-1])
    out = np.zeros(f.usize - 5, dtype=np.uint8)
    with pytest.raises(hb.HuffError) as e:
        m.decode_host(f.tree, f.data, f.bits, out)
    assert e.value.code == -6
    m.close()


@pytest.mark.parametrize("kind,log2n", [(0, 24), (1, 25), (2, 22)])
def test_multi_resident_generate_decode_verify(kind, log2n):
    """one synthetic stream built on the devices, decoded resident, every slice verified"""
    for n in _counts():
        m = hb.Multi(n)
        bits = m.generate(kind, SEED, 1 << log2n)
        assert bits > 0
        for _ in range(2):
            res = m.decode()
            assert res["n_symbols"] == 1 << log2n, (n, res)
        assert m.verify(kind, SEED) == 0
        # and the bytes themselves against the CPU generator
        out = np.zeros(1 << log2n, dtype=np.uint8)
        m.download(out)
        want = hb.Model(kind).symbols_cpu(SEED, 0, 1 << log2n)
        assert np.array_equal(out, want), n
        m.close()


def test_multi_load_truncated_stream():
    """a stream cut inside a codeword: the cut-off codeword emits nothing, on any split"""
    f = _stream("paper1")
    st = O.load_huff(O.corpus_path("paper1"))
    for bits in (f.bits - 1, f.bits - 7, 100001):
        want = O.simple_decode(st, bits=bits)
        for n in _counts():
            m = hb.Multi(n)
            out = np.zeros(want.size + 8, dtype=np.uint8)
            res = m.decode_host(f.tree, f.data, bits, out[: want.size])
            assert res["n_symbols"] == want.size, (bits, n)
            assert np.array_equal(out[: want.size], want)
            m.close()


def test_onethread_matches_oracle():
    """the reference's onethread debug approach (framework/onethread.cu:13-52)"""
    for name in ("hello", "paper1"):
        f = _stream(name)
        ctx = hb.Context(0)
        out = np.zeros(f.usize, dtype=np.uint8)
        res = hb.decode_onethread(ctx, f.tree, f.data, f.bits, out)
        assert res["n_symbols"] == f.usize
        assert O.sha256(out) == O.CORPORA[name][2]
        ctx.close()
