"""torchrun worker: multi-GPU run == single-GPU run, bit for bit.
Launched by tests/test_gpu_multi.py as
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_worker.py
Each rank owns one GPU, one METIS/slab subdomain and exchanges halo cells over NCCL send/recv."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    os.environ.setdefault("CFD2D_TILE", "128")      # several interior + boundary tiles per rank on this small mesh
    os.environ.setdefault("CFD2D_PIPE_TILE", "96")
    import torch
    import torch.distributed as dist
    from cfd2d_b200 import cases, decomp, fvm
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # step layouts of a multi-rank handle: three sweeps with the halo exchange overlapped on the comm
    # stream (default), the same serialised on one stream, the tile-fused stage kernel, and the
    # pipelined tile kernel (interior tiles under the exchanges, boundary tiles after them)
    # "graph": the overlapped step captured into one two-stream CUDA graph (the default launches it eagerly)
    modes = {"overlap": {"CFD2D_FUSED": "0", "CFD2D_OVERLAP": "1", "CFD2D_GRAPH_MULTI": "0", "CFD2D_HALO_P2P": "0"},
             "graph": {"CFD2D_FUSED": "0", "CFD2D_OVERLAP": "1", "CFD2D_GRAPH_MULTI": "1", "CFD2D_HALO_P2P": "0"},
             "serial": {"CFD2D_FUSED": "0", "CFD2D_OVERLAP": "0", "CFD2D_GRAPH_MULTI": "0", "CFD2D_HALO_P2P": "0"},
             "fused": {"CFD2D_FUSED": "1", "CFD2D_OVERLAP": "1", "CFD2D_GRAPH_MULTI": "1", "CFD2D_HALO_P2P": "0"},
             "pipe": {"CFD2D_FUSED": "2", "CFD2D_OVERLAP": "1", "CFD2D_GRAPH_MULTI": "0", "CFD2D_HALO_P2P": "0"},
             # halo records stored straight into the neighbours' arrays over NVLink (falls back to NCCL where a
             # neighbour cannot be mapped; the line printed below names the transport that ran)
             "p2p": {"CFD2D_FUSED": "0", "CFD2D_OVERLAP": "1", "CFD2D_GRAPH_MULTI": "0", "CFD2D_HALO_P2P": "1"},
             "p2p_graph": {"CFD2D_FUSED": "0", "CFD2D_OVERLAP": "1", "CFD2D_GRAPH_MULTI": "1", "CFD2D_HALO_P2P": "1"}}
    if os.environ.get("WORKER_MODES"):               # subset, e.g. WORKER_MODES=p2p,p2p_graph
        modes = {k: v for k, v in modes.items() if k in os.environ["WORKER_MODES"].split(",")}
    # (partition, flux, order, steady, p_max): the last two cases lower the pressure limit below the
    # initial peak so that a blob of adjacent cells trips the limiter ACROSS the partition cut:
    # remediateLimCells (fvm_tvd.cpp:464-499) then rewrites send cells after the end-of-step exchange
    # and reads flagged halo neighbours -- the per-rank sweep must still equal the serial one
    cases_ = (("metis", 0, 2, 0, None), ("slab", 1, 2, 0, None), ("metis", 1, 1, 0, None), ("slab", 0, 2, 1, None),
              ("slab", 1, 2, 0, 1.005e5), ("metis", 0, 2, 0, 1.005e5))
    for mode, partition, flux, order, steady, p_max in [(m,) + c for m in modes for c in cases_]:
        os.environ.update(modes[mode])
        c = cases.channel(96, 48, jitter=0.2, shuffle=True)
        c.task.steady = steady
        if p_max is not None:
            c.task.p_max = p_max
        st = c.smooth_state()
        if partition == "metis" and not os.path.exists(decomp.METIS_LIB):
            partition = "slab"
        s, st_loc, nc_loc, nc_tot = decomp.make_rank_solver(0, 0, rank, world, local, flux, order, dist,
                                                            partition=partition, case=c, state=st)
        s.set_state(*st_loc)
        tau = s.calc_time_step()
        s.step(12)
        got = s.get_state()
        transport = s.halo_transport
        rm = s.rank_mesh
        nflag = torch.zeros(world, dtype=torch.int64, device="cuda")
        nflag[rank] = int((got[5] != 0).sum())           # cells the limiter touched on this rank
        dist.all_reduce(nflag)
        # assemble the global state on every rank
        glob = [torch.zeros(nc_tot, dtype=torch.float64, device="cuda") for _ in range(4)]
        idx = torch.from_numpy(rm.g_cells[:rm.nc].astype(np.int64)).cuda()
        for k in range(4):
            glob[k][idx] = torch.from_numpy(got[k]).cuda()
            dist.all_reduce(glob[k])
        s.close()
        if rank == 0:
            os.environ["CFD2D_FUSED"] = "0"
            s1 = fvm.Solver(c.mesh, c.task, flux, order, device=local)
            s1.set_state(*st)
            tau1 = s1.calc_time_step()
            s1.step(12)
            ref = s1.get_state()
            s1.close()
            same = all(np.array_equal(glob[k].cpu().numpy(), ref[k]) for k in range(4)) and tau == tau1
            extra = ""
            if p_max is not None:
                # the limiter must really have tripped on both sides of a cut, and the run must still be
                # the reference's: C oracle (serial ascending sweep), LF bit-exact / Godunov <= 1e-12
                from oracle import port
                import parity_cases as pc
                ranks_hit = int((nflag > 0).sum().item())
                o = port.OracleSolver(c.mesh, c.task, flux, order)
                o.set_state(*st); o.calc_time_step(); o.step(12)
                oref = o.get_state(); o.close()
                err = max(pc.err_norm(ref[:4], oref[:4]))
                flags_ok = bool(np.array_equal(ref[5], oref[5]))
                good = ranks_hit >= 2 and flags_ok and (err == 0.0 if flux == 1 else err < 1e-12)
                extra = f" limiter: touched cells per rank={nflag.tolist()} err_vs_oracle={err:.2e} flags_equal={flags_ok} ok={good}"
                same = same and good
            print(f"[multi-gpu x{world}] mode={mode} ({transport}) partition={partition} flux={flux} order={order} steady={steady} p_max={p_max}: "
                  f"bitwise equal to 1 GPU = {same}, tau={tau}{extra}", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
