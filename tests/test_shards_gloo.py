"""World-size-2 (and 3) run of the multi-GPU driver logic on CPU with the gloo
backend: byte-range sharding, one all_gather of the 32-entry maps, composition,
independent emit.  The per-shard kernels are replaced by their CPU emulation
(tests/emul); the exchange/compose logic is the one bench.py runs on NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import emul_lib as E
    import oracle_lib as O
    import huffmandecoderongpus_b200 as hb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = O.load_huff(O.corpus_path(name))
    lut = hb.build_lut(st.tree)
    per = (st.nbytes // world) // 16 * 16
    a = rank * per
    b = st.nbytes if rank == world - 1 else (rank + 1) * per
    last = rank == world - 1
    bits_own = st.bits - 8 * a if last else 8 * (b - a)
    halo_end = min(st.nbytes, b + 16)
    bits_avail = bits_own if last else min(st.bits - 8 * a, 8 * (halo_end - a))
    words = E.words_of(st.data[a:], halo_end - a)
    _, smap, _, _, rc = E.run(lut, words, bits_own, bits_avail, emit=False)
    assert rc == 0
    mine = torch.from_numpy(smap.astype(np.int64))
    allm = torch.zeros(32 * world, dtype=torch.int64)
    dist.all_gather_into_tensor(allm, mine)
    maps = allm.numpy().astype(np.uint64).reshape(world, 32)
    cur, base = 0, 0                      # what hb_compose_kernel does on the device
    for r in range(rank):
        m = int(maps[r][cur]); base += m >> 8; cur = m & 31
    out, _, res, _, rc = E.run(lut, words, bits_own, bits_avail, entry=cur, base=base)
    assert rc == 0
    n = int(res[0])
    tot = torch.tensor([n], dtype=torch.int64)
    dist.all_reduce(tot)
    want = O.simple_decode(st)
    ok = int(tot[0]) == want.size and np.array_equal(out[:n], want[base: base + n])
    q.put((rank, ok, base, n))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "paper1"), (3, "news")])
def test_two_rank_exchange(world, name):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    if O.corpus_path(name) is None:
        pytest.skip("corpus not present")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == 0 and all(res[i][2] == res[i - 1][2] + res[i - 1][3] for i in range(1, world))
