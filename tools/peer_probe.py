"""Peer-memory all-reduce (csrc/peer.cu) against torch.distributed's NCCL all-reduce, under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/peer_probe.py
Checks the sums (bit-identical across ranks, equal to NCCL's), a sharded Sobol analysis end to end, and times both."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import gsa_loader
g = gsa_loader.load()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = g.default_handle(local)
out = {"world": world}

for nboot, nout, ns in ((1, 1, 208), (1000, 1, 17), (1, 256, 48)):
    pg = g.PeerGroup(h, nboot * nout * ns)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.randn((nboot, nout, ns), dtype=torch.float64, device="cuda", generator=gen)
    x[..., 1:8:2] *= 1e-17                                  # lo words of the double-double entries
    ref = x.clone()
    dist.all_reduce(ref)
    mine = x.clone()
    host = torch.empty(x.shape, dtype=torch.float64).pin_memory()
    pg.allreduce(mine, ns, host_out=host)
    torch.cuda.synchronize()
    assert pg.status() == 0
    hi_lo = lambda t: t[..., 0:8:2] + t[..., 1:8:2]
    err_plain = float((mine[..., 8:] - ref[..., 8:]).abs().max())
    err_dd = float((hi_lo(mine) - hi_lo(ref)).abs().max())
    assert torch.equal(mine.cpu(), host), "host copy differs from the device result"
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    assert all(torch.equal(allr[0], t) for t in allr), "ranks disagree"
    assert err_plain <= 1e-12 and err_dd <= 1e-12, (err_plain, err_dd)
    # timing: exchange + result on the host, per call
    reps = 300
    for _ in range(20):
        pg.allreduce(mine, ns, host_out=host); torch.cuda.current_stream().synchronize()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        pg.allreduce(mine, ns, host_out=host)
        torch.cuda.current_stream().synchronize()
    t_peer = (time.perf_counter() - t0) / reps
    y = x.clone()
    for _ in range(20):
        dist.all_reduce(y); host.copy_(y, non_blocking=True); torch.cuda.current_stream().synchronize()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dist.all_reduce(y)
        host.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t_nccl = (time.perf_counter() - t0) / reps
    out[f"{nboot}x{nout}x{ns}"] = {"peer_us": round(1e6 * t_peer, 1), "nccl_us": round(1e6 * t_nccl, 1), "err_plain": err_plain, "err_dd": err_dd}
    pg.close()

# end to end: sharded config-5 shape (small N) through the peer all-reduce equals the single-GPU analysis
d, N = 100, 1 << 18
a = [0.0, 1.0, 4.5, 9.0] + [99.0] * 96
model, method = g.DeviceModel("sobol_g", a), g.Sobol(nboot=3)
ns = g.nstat(method, d)
pg = g.PeerGroup(h, 3 * ns)
c0, nc = g.shard_columns(N, rank, world)
A, B = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, col0=c0, ncols=nc, handle=h)
p = g.sobol_partials(model, method, A, B, N, c0, handle=h)
host = torch.empty(p.shape, dtype=torch.float64).pin_memory()
pg.allreduce(p, ns, host_out=host)
torch.cuda.current_stream().synchronize()
res = g.sobol_finalize(method, host.numpy(), d, N)
Aw, Bw = g.generate_design_matrices(N, np.zeros(d), np.ones(d), 2, device=True, handle=h)
whole = g.gsa(model, method, Aw, Bw, batch=True, handle=h)
err = max(float(np.abs(res.S1 - whole.S1).max()), float(np.abs(res.ST - whole.ST).max()), float(np.abs(res.ST_Conf_Int - whole.ST_Conf_Int).max()))
assert err <= 1e-12, err
out["e2e_max_abs_diff_vs_single_gpu"] = err
# the whole step as ONE call / ONE kernel (gsa_sobol_run_sharded: the fused kernel's last CTA reduces the tiles, exchanges the
# sums over NVLink and writes them to pinned host memory), with a resident and with a generated design; second order (no fused
# tail: the classic launches follow) as well
step = g.ShardedStep(model, method, A, B, N, c0, handle=h)
for _ in range(3):
    r1 = step.run()
launches = h.last_timing()[2]
err = max(float(np.abs(r1.S1 - whole.S1).max()), float(np.abs(r1.ST - whole.ST).max()), float(np.abs(r1.ST_Conf_Int - whole.ST_Conf_Int).max()))
assert err <= 1e-12, ("fused sharded step", err)
out["sharded_step_max_abs_diff_vs_single_gpu"] = err
out["sharded_step_launches"] = launches
allr = [None] * world
dist.all_gather_object(allr, (r1.S1.tobytes(), r1.ST.tobytes(), r1.ST_Conf_Int.tobytes()))
assert all(x == allr[0] for x in allr), "ranks disagree on the fused sharded step"
# (generated: N = nboot * samples columns, replicate r uses its own dimension block -- compare with the single-GPU generated run)
gw = g.gsa(model, method, [[0.0, 1.0]] * d, samples=N // 3, batch=True, handle=h)
sg = g.ShardedStep(model, method, None, None, (N // 3) * 3, *g.shard_columns((N // 3) * 3, rank, world)[:1], handle=h,
                   p_range=(np.zeros(d), np.ones(d)), ncols=g.shard_columns((N // 3) * 3, rank, world)[1])
r2 = sg.run()
err = max(float(np.abs(r2.S1 - gw.S1).max()), float(np.abs(r2.ST - gw.ST).max()))
assert err <= 1e-12, ("generated sharded step", err)
out["sharded_generated_step_max_abs_diff_vs_single_gpu"] = err
a8 = [0.0, 1.0, 4.5, 9.0, 99.0, 99.0, 99.0, 99.0]
m8, me8 = g.DeviceModel("sobol_g", a8), g.Sobol(order=[0, 1, 2], nboot=2)
pg.close()
pg = g.PeerGroup(h, 2 * g.nstat(me8, 8))
N8 = 1 << 16
c8, n8 = g.shard_columns(N8, rank, world)
A8, B8 = g.generate_design_matrices(N8, np.zeros(8), np.ones(8), 2, device=True, handle=h)
w8 = g.gsa(m8, me8, A8, B8, batch=True, handle=h)
r8 = g.ShardedStep(m8, me8, A8[:, c8:c8 + n8], B8[:, c8:c8 + n8], N8, c8, handle=h).run()
err = max(float(np.abs(r8.S1 - w8.S1).max()), float(np.abs(r8.S2 - w8.S2).max()), float(np.abs(r8.S2_Conf_Int - w8.S2_Conf_Int).max()))
assert err <= 1e-12, ("unfused sharded step", err)
out["sharded_step_second_order_max_abs_diff"] = err
# timing of the one-call step against partials + allreduce + finalize (small shard: the overhead is what shows)
step = g.ShardedStep(model, method, A, B, N, c0, handle=h)
pg.close()
pg = g.PeerGroup(h, 3 * ns)
for _ in range(10):
    step.run()
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    step.run()
t_step = (time.perf_counter() - t0) / 100
kms = h.last_timing()[1]
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    p = g.sobol_partials(model, method, A, B, N, c0, handle=h, out=p)
    pg.allreduce(p, ns, host_out=host)
    torch.cuda.current_stream().synchronize()
    g.sobol_finalize(method, host.numpy(), d, N)
t_old = (time.perf_counter() - t0) / 100
out["step_us"] = {"one_call_fused": round(1e6 * t_step, 1), "three_calls": round(1e6 * t_old, 1), "kernel_us": round(1e3 * kms, 1)}
h.set_stream(None)
pg.close()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
