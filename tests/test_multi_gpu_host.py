"""world_size-2 gloo test (CPU) of the multi-GPU host logic: each rank packs its shard of the board into partial
digests, the partial digests are summed with all_reduce and reduced mod q2 — the result must equal the unsharded
digest.  The shard packing itself is done by the oracle here (no GPU in this container); the sharding,
global-index bookkeeping and the reduction are the code under test."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O

Q2 = O.Q2


def shard_bounds(D, world, rank):
    """static equal split, remainder to the first D % world ranks"""
    base, rem = divmod(D, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _worker(rank, world, port, D, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = np.load(path)
    pv, payloads, weights = data["pv"], data["payloads"], data["weights"]
    lo, hi = shard_bounds(D, world, rank)
    n_idx, n_pay = int(data["n_idx"]), int(data["n_pay"])
    part = np.zeros((n_idx + n_pay, 2, 2048), np.uint64)
    for c in range(n_idx):
        part[c] = O.encode_indices(D, 2, pv[lo:hi], lo, 99, c)
    part[n_idx:] = O.encode_payloads(pv[lo:hi], payloads[lo:hi], lo, weights, n_pay, threads=1)
    t = torch.from_numpy(part.view(np.int64).copy())
    dist.all_reduce(t)                                            # integer sum: world * q2 < 2^63
    total = (t.numpy().view(np.uint64) % np.uint64(Q2)).astype(np.uint64)
    if rank == 0:
        np.save(path + ".out.npy", total)
    dist.destroy_process_group()


def test_sharded_digest_sum_equals_unsharded(tmp_path):
    D, world = 7, 2
    rng = np.random.default_rng(1)
    pv = rng.integers(0, Q2, (D, 2, 2048), dtype=np.uint64)         # arbitrary NTT-domain ciphertexts: packing is linear
    payloads = rng.integers(0, 256, (D, O.PAYLOAD_LEN), dtype=np.uint16)
    rp = O.retrieval_params(D, 2)
    n_idx, n_pay = rp["max_encode_indices_cipher_count"], rp["payload_cipher_count"]
    weights = rng.integers(0, 257, (n_pay * 2, D), dtype=np.uint16)
    path = str(tmp_path / "in.npz")
    np.savez(path, pv=pv, payloads=payloads, weights=weights, n_idx=n_idx, n_pay=n_pay)
    assert [shard_bounds(D, world, r) for r in range(world)] == [(0, 4), (4, 7)]
    mp.spawn(_worker, args=(world, 29517, D, path), nprocs=world, join=True)
    got = np.load(path + ".out.npy")
    ref = np.zeros_like(got)
    for c in range(n_idx):
        ref[c] = O.encode_indices(D, 2, pv, 0, 99, c)
    ref[n_idx:] = O.encode_payloads(pv, payloads, 0, weights, n_pay, threads=1)
    assert np.array_equal(got, ref)
