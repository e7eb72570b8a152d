"""The N > 1 path: particles sharded over ranks with a GLOBAL systematic resample.

CPU (gloo, world_size 2): the sharding protocol -- block statistics per rank, all-reduce of the
block exponent, all-gather of rank records (total + prefix at the margin edges), the margin check
every rank evaluates for every rank, and the fixed-margin halo exchange with the two neighbours --
restated with exact Python integers over torch.distributed, must reproduce the single-process
oracle's ancestors bit for bit.

GPU (-m gpu, needs >= 2 devices): the CUDA/NCCL implementation on 2 ranks equals the
single-GPU run at the same seed bit for bit (counter-based RNG keyed on the global slot)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import pzhelpers as H

BLK = 256


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


GUARD = 16
MARGIN = 8   # halo margin in blocks of the protocol test (16 blocks per rank)


def _block_weights(lw32):
    """q (16-bit weights at the block scale) and block exponents from the oracle's spec functions."""
    q, _, eb = H.block_weights(np.ascontiguousarray(lw32, dtype=np.float32))
    return q.astype(np.int64), eb.astype(np.int64)


def _shift_guard(s, ksh):
    return (s << (GUARD - ksh)) if ksh <= GUARD else (s >> (ksh - GUARD))


def _block_pos(Cv, rho, L, U):
    return (((int(Cv) << (64 - L)) * rho) >> 64) + (0xFFFFFFFF - U)


def _slot_end(Cv, rho, L, U, n_out):
    return min(_block_pos(Cv, rho, L, U) >> 32, n_out)


def _cdf_offset(p):
    return ((p >> 2) & 1) * 128 + (p >> 3) * 4 + (p & 3)


def _gloo_worker(rank, world, port, n_global, seed, sigma, shift, ret):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pz = g.load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    lw = (rng.standard_normal(n_global) * sigma).astype(np.float32)
    lw[rng.random(n_global) < 0.05] = -np.inf
    lw[: n_global // 3] -= shift   # unbalanced rank totals => real migration
    U = 0x12345678
    n_local = n_global // world
    nb = n_local // BLK
    mine = lw[rank * n_local:(rank + 1) * n_local]
    NZ = -(1 << 30)
    # per-block statistics (what the step kernel writes): 16-bit weights, block exponent, block sum
    q_loc, eb = _block_weights(mine)
    sb = [int(q_loc[b * BLK:(b + 1) * BLK].sum()) for b in range(nb)]
    # (1) all-reduce max of the block exponent
    e = torch.tensor([int(eb.max())]); dist.all_reduce(e, op=dist.ReduceOp.MAX); emax = int(e.item())
    # (2) local scan at the global exponent, all-gather of rank totals
    S = [_shift_guard(sb[b], int(emax - eb[b])) if eb[b] != NZ else 0 for b in range(nb)]
    pb = [0]
    for v in S:
        pb.append(pb[-1] + v)
    tot = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(tot, torch.tensor([pb[-1]], dtype=torch.int64))
    totals = [int(t.item()) for t in tot]
    base, T = sum(totals[:rank]), sum(totals)
    L = T.bit_length(); rho = ((n_global << (32 + L)) // T) + 1
    lam = rho.bit_length() - 30; Mq = (rho >> lam) + 1; h0 = L - GUARD - lam
    pbx = [base + v for v in pb]
    # (3) margin check from the gathered records alone (every rank reaches the same verdict):
    #     rank q sends its first M blocks to q-1 and its last M blocks to q+1
    M = MARGIN
    rec = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(rec, torch.tensor([pb[-1], pb[min(M, nb)], pb[nb - min(M, nb)]], dtype=torch.int64))
    overflow, bq = False, 0
    for qi in range(world):
        tq, p_first, p_last = (int(v) for v in rec[qi])
        if qi > 0 and _slot_end(bq + p_first, rho, L, U, n_global) < qi * n_local:
            overflow = True
        if qi + 1 < world and _slot_end(bq + p_last, rho, L, U, n_global) > (qi + 1) * n_local:
            overflow = True
        bq += tq
    assert not overflow
    s0, s1 = _slot_end(pbx[0], rho, L, U, n_global), _slot_end(pbx[nb], rho, L, U, n_global)
    j_lo, j_hi = rank * n_local, (rank + 1) * n_local
    migrated = max(0, min(s0, j_hi) - j_lo) + max(0, j_hi - max(s1, j_lo))
    # (4) fixed-margin halo exchange with the two neighbours (weights, block exponents, absolute prefixes)
    ext = {}  # global block -> (q[256], eb, pbx_start, pbx_end)
    for b in range(nb):
        ext[rank * nb + b] = (q_loc[b * BLK:(b + 1) * BLK], int(eb[b]), pbx[b], pbx[b + 1])

    def pack(blocks):
        return torch.tensor(np.concatenate([np.concatenate([ext[gb][0].astype(np.float64), [ext[gb][1], ext[gb][2], ext[gb][3]]])
                                            for gb in blocks]))
    reqs, recv_bufs = [], []
    for peer, mine_blocks, theirs in ((rank - 1, range(rank * nb, rank * nb + M), range(rank * nb - M, rank * nb)),
                                      (rank + 1, range((rank + 1) * nb - M, (rank + 1) * nb), range((rank + 1) * nb, (rank + 1) * nb + M))):
        if peer < 0 or peer >= world:
            continue
        reqs.append(dist.isend(pack(mine_blocks), peer))
        buf = torch.zeros(M * (BLK + 3), dtype=torch.float64)
        reqs.append(dist.irecv(buf, peer)); recv_bufs.append((theirs, buf))
    for r in reqs:
        r.wait()
    for blocks, buf in recv_bufs:
        a = buf.numpy().reshape(M, BLK + 3)
        for k, gb in enumerate(blocks):
            ext[gb] = (a[k, :BLK].astype(np.int64), int(a[k, BLK]), int(a[k, BLK + 1]), int(a[k, BLK + 2]))
    need_lo, need_hi = min(ext), max(ext)
    # (5) resample my output slots from the (halo-extended) blocks
    anc = np.full(n_local, -1, dtype=np.int64)
    for gb in range(need_lo, need_hi + 1):
        bq, ebb, P0, P1 = ext[gb]
        X0, X1 = _block_pos(P0, rho, L, U), _block_pos(P1, rho, L, U)
        o0, o1 = X0 >> 32, X1 >> 32
        if o1 == o0:
            continue
        h = h0 + (emax - ebb)
        A = X1 & 0xFFFFFFFF
        sbb = int(bq.sum())
        c = 0
        prev = o0
        for p in range(BLK):  # the block's CDF order
            i = _cdf_offset(p)
            c += int(bq[i])
            r = sbb - c
            num = (A << h) - r * Mq if h >= 0 else A - ((r * Mq) << (-h))   # floor((A 2^h - r M) / 2^(h+32))
            o = max(o1 + (num >> (max(h, 0) + 32)), o0)
            for j in range(max(prev, j_lo), min(o, j_hi)):
                anc[j - j_lo] = gb * BLK + i
            prev = max(prev, o)
    out = [torch.zeros(n_local, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, torch.from_numpy(anc))
    if rank == 0:
        full = np.concatenate([t.numpy() for t in out])
        rc, ref = H.oracle_systematic(lw, U)
        ret["ok"] = bool(rc == 0 and np.array_equal(full, ref))
        ret["mismatch"] = int(np.sum(full != ref))
    ret["migrated%d" % rank] = migrated
    dist.destroy_process_group()


def test_sharding_protocol_gloo_world2():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, 8192, 17, 3.0, 6.0, ret), nprocs=2, join=True)
    assert ret.get("ok"), dict(ret)
    assert ret["migrated0"] > 0  # the skewed weights make rank 0 take ancestors from rank 1's margin


@pytest.mark.parametrize("shift", [0.25, -0.25])
def test_sharding_protocol_gloo_world4_middle_ranks(shift):
    """Four ranks: the two middle ranks exchange margins with BOTH neighbours.  Tilting the weights towards the
    later (shift > 0) or the earlier (shift < 0) particles moves ancestry across every boundary to the right or to
    the left, so both halos of the middle ranks are read (the tilt is mild enough for the 8-block margin)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_gloo_worker, args=(4, port, 16384, 23, 0.5, shift, ret), nprocs=4, join=True)
    assert ret.get("ok"), dict(ret)
    takers = range(3) if shift > 0 else range(1, 4)   # ranks whose slots reach into a neighbour's margin
    assert all(ret["migrated%d" % r] > 0 for r in takers), dict(ret)


def _gpu_worker(rank, world, port, model, n_global, seed, T, env, outdir):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    os.environ.update(env or {})
    pz = g.load_package()
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ids = [pz.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, 0)
    d = getattr(pz, model)()
    hd = {"rw1d": H.model_rw1d, "mtt_track": H.model_mtt_track, "stt": H.model_stt}[model]()
    obs, _ = H.gen_obs(hd, T)
    zs = pz.zparticles_sharded(n_global, d, ids[0], rank, world, seed=seed, device=rank)
    if env and env.get("PZ_HALO_BLOCKS") == "1" and env.get("PZ_SHARD_NCCL"):
        # a one-block halo margin cannot hold the migration of this stream: on the NCCL exchange path every rank reports it
        codes = []
        for y in obs:
            try:
                zs.step(y); codes.append(0)
            except pz.PzError as e:
                codes.append(e.code)
        np.save(os.path.join(outdir, f"codes_{rank}.npy"), np.array(codes))
        zs.close(); dist.destroy_process_group()
        return
    posts = [zs.step(y) for y in obs]
    np.save(os.path.join(outdir, f"x_{rank}.npy"), zs.read_state())
    np.save(os.path.join(outdir, f"s_{rank}.npy"), np.array([[p.total_weight, p.e_max, p.n_migrated] + list(p.mean) for p in posts], dtype=np.float64))
    zs.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["peer_memory", "nccl"])
@pytest.mark.parametrize("model", ["rw1d", "mtt_track", "stt"])
def test_sharded_equals_single_gpu(pz, model, exchange, world, tmp_path):
    """Both exchange implementations (kernels over NVLink peer memory; NCCL calls between the kernels),
    at every rank count the box offers."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n_global, seed, T = world * 256 * 101, 99, 6  # partial last tile on every rank
    env = {"PZ_SHARD_NCCL": "1"} if exchange == "nccl" else {}
    mp.spawn(_gpu_worker, args=(world, _free_port(), model, n_global, seed, T, env, str(tmp_path)), nprocs=world, join=True)
    d = getattr(pz, model)()
    hd = {"rw1d": H.model_rw1d, "mtt_track": H.model_mtt_track, "stt": H.model_stt}[model]()
    obs, _ = H.gen_obs(hd, T)
    zs = pz.zparticles(n_global, d, seed=seed)
    posts = [zs.step(y) for y in obs]
    x = zs.read_state()
    xs = np.concatenate([np.load(tmp_path / f"x_{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(xs.view(np.uint32), x.view(np.uint32))
    s0 = np.load(tmp_path / "s_0.npy")
    for t, p in enumerate(posts):
        assert int(s0[t, 0]) == p.total_weight and int(s0[t, 1]) == p.e_max
        assert np.allclose(s0[t, 3:3 + d.nx], p.mean, rtol=1e-6, atol=2e-6)


@pytest.mark.gpu
def test_sharded_margin_overflow_nccl_path_reports_it_on_every_rank(pz, tmp_path):
    """PZ_HALO_BLOCKS=1 on the NCCL exchange path (kept for A/B): the ancestors of a rank's boundary slots reach beyond a
    one-block margin within a few steps; the verdict is computed by every rank from the gathered records, so both ranks
    raise PZ_ECAPACITY (-5) at the same step instead of resampling from an incomplete halo."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n_global, seed, T = 2, 2 * 256 * 400, 7, 40
    mp.spawn(_gpu_worker, args=(world, _free_port(), "rw1d", n_global, seed, T, {"PZ_HALO_BLOCKS": "1", "PZ_SHARD_NCCL": "1"}, str(tmp_path)),
             nprocs=world, join=True)
    c0, c1 = np.load(tmp_path / "codes_0.npy"), np.load(tmp_path / "codes_1.npy")
    assert np.array_equal(c0, c1)
    assert (c0 != 0).any() and c0[np.flatnonzero(c0)[0]] == -5   # the first failure is the capacity report


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["rw1d", "stt"])
def test_sharded_margin_overflow_falls_back_to_remote_reads(pz, model, tmp_path):
    """The peer-memory path never fails on imbalance (the reference's resample has no such failure mode,
    Inference.hs:60-61): with a one-block margin (PZ_HALO_BLOCKS=1) the ancestors of boundary slots reach beyond it
    within a few steps; those steps run the remote-read kernels (global merge-path split, blocks read from their
    owners over NVLink) and the run stays bit-equal to the single-GPU filter."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    n_global, seed, T = world * 256 * 400, 7, 40
    mp.spawn(_gpu_worker, args=(world, _free_port(), model, n_global, seed, T, {"PZ_HALO_BLOCKS": "1"}, str(tmp_path)), nprocs=world, join=True)
    d = getattr(pz, model)()
    hd = {"rw1d": H.model_rw1d, "stt": H.model_stt}[model]()
    obs, _ = H.gen_obs(hd, T)
    zs = pz.zparticles(n_global, d, seed=seed)
    posts = [zs.step(y) for y in obs]
    xs = np.concatenate([np.load(tmp_path / f"x_{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(xs.view(np.uint32), zs.read_state().view(np.uint32))
    s0 = np.load(tmp_path / "s_0.npy")
    assert [int(v) for v in s0[:, 0]] == [p.total_weight for p in posts]
    assert s0[:, 2].min() >= 0 and s0[:, 2].max() > 256      # migration counts, some beyond one block: the margin did overflow


@pytest.mark.gpu
@pytest.mark.parametrize("mk,f64", [("mtt_track", False), ("mtt_track", True), ("stt", True), ("rw1d", True)],
                         ids=["lg42-f32", "lg42-f64", "lg63-f64", "rw1d-f64"])
def test_multi_device_margin_overflow_falls_back_to_remote_reads(pz, monkeypatch, mk, f64):
    """The same through pz_filter_create_multi (one process), on a stream with outliers that unbalances the ranks; every
    storage type whose step kernel stages parent INDICES (the remote kernel then stages global indices and gathers the
    parent from its owner) and the state-staged fp32 path."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("PZ_HALO_BLOCKS", "1")
    hd = {"mtt_track": H.model_mtt_track, "stt": H.model_stt, "rw1d": H.model_rw1d}[mk]()
    T = 30
    obs, _ = H.gen_obs(hd, T)
    obs = obs.copy(); obs[5] += 6.0; obs[17] -= 9.0
    world = 2 if ndev < 4 else 4
    n = world * 256 * 300
    one = pz.zparticles(n, getattr(pz, mk)(f64=f64), seed=5)
    many = pz.zparticles_multi(n, getattr(pz, mk)(f64=f64), list(range(world)), seed=5)
    mig = 0
    for t in range(T):
        pa, pb = one.step(obs[t]), many.step(obs[t])
        assert pa.total_weight == pb.total_weight and pb.n_migrated >= 0, t
        mig = max(mig, pb.n_migrated)
        assert np.array_equal(one.read_state().view(np.uint32), many.read_state().view(np.uint32)), t
    assert mig > 256 * world
    one.close(); many.close()


@pytest.mark.gpu
@pytest.mark.parametrize("f64", [False, True], ids=["f32", "f64"])
@pytest.mark.parametrize("model", ["rw1d", "mtt_track", "stt"])
def test_multi_device_single_process_equals_single_gpu(pz, model, f64):
    """pz_filter_create_multi (one host thread, every device of the box): the same shards wired with
    direct peer pointers and driven from this process; bit-equal to the single-GPU filter, per step and over a
    pz_filter_run batch, at every device count the box offers."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    hd = {"rw1d": H.model_rw1d, "mtt_track": H.model_mtt_track, "stt": H.model_stt}[model]()
    T = 8
    obs, _ = H.gen_obs(hd, T)
    bt = np.uint64 if f64 else np.uint32
    for world in [w for w in (2, 4, 8) if w <= ndev]:
        n = world * 256 * 101
        d = getattr(pz, model)(f64=f64)
        one = pz.zparticles(n, d, seed=99)
        many = pz.zparticles_multi(n, d, list(range(world)), seed=99)
        assert many.n == n
        for t in range(4):
            pa, pb = one.step(obs[t]), many.step(obs[t])
            assert pa.total_weight == pb.total_weight and pa.e_max == pb.e_max and pb.step == t and pb.n_migrated >= 0
            assert np.allclose(pa.mean, pb.mean, rtol=1e-6, atol=2e-6)
            assert np.array_equal(one.read_state().view(bt), many.read_state().view(bt)), (model, world, t)
        pa, pb = one.run(obs[4:]), many.run(obs[4:])
        assert [p.total_weight for p in pa] == [p.total_weight for p in pb]
        assert np.array_equal(one.read_state().view(bt), many.read_state().view(bt))
        ms, launches = many.last_timing()
        assert ms > 0 and launches == 4 * world * 4   # step kernel, its remote-read twin (exits at once), scan + margin push, plan + partition
        with pytest.raises(pz.PzError):
            many.read_ancestors()
        one.close(); many.close()
