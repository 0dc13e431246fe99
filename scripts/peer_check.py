"""N-GPU check of the peer-memory gradient exchange (egn_peer_allreduce) against ncclAllReduce, under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/peer_check.py [--out f.json]

1. raw buffers of the training size (24.8 M floats): peer sum == NCCL sum (bit-identical at N = 2, 1e-6 relative otherwise),
   identical on every rank; device time per call of both (CUDA events, max over ranks) for several grid sizes;
2. three TableAdam training steps of a 128^3 scene from identical initial state with either exchange: the factor tables agree
   (atomics make a step non-deterministic at the 1e-6 level, so the comparison is relative).
Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--numel", type=int, default=24721124 + 56320)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from egonerf_b200.sharding import PeerExchange
    from egonerf_b200.optim import TableAdam
    from egonerf_b200.scene_io import model_from_scene, RENDER_KW
    from egonerf_b200.synthetic import make_scene, make_rays

    def maxr(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"world": world, "numel": args.numel}
    # ---- 1. raw buffers ----
    px = PeerExchange(args.numel, dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(args.numel, device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref)
    px.tensor[:args.numel].copy_(src)
    px.allreduce(1.0)
    torch.cuda.synchronize()
    got = px.tensor[:args.numel]
    res["max_abs_diff_vs_nccl"] = maxr((got - ref).abs().max().item())
    res["bit_identical_to_nccl"] = bool(maxr(0.0 if torch.equal(got, ref) else 1.0) == 0.0)
    chk = got.double().sum().item()
    lo, hi = torch.tensor([chk], device=dev, dtype=torch.float64), torch.tensor([chk], device=dev, dtype=torch.float64)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["identical_on_all_ranks"] = bool(lo.item() == hi.item())

    def timeit(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return maxr(a.elapsed_time(b) / args.reps)

    res["nccl_ms"] = timeit(lambda: dist.all_reduce(ref))
    res["peer_ms"] = {}
    for blocks in (37, 74, 148, 256):
        px.blocks = blocks
        res["peer_ms"][str(blocks)] = timeit(lambda: px.allreduce(1.0))
    res["mb"] = args.numel * 4 / 1e6
    px.close()

    # ---- 2. training steps with either exchange ----
    def run(exchange):
        torch.manual_seed(1)
        scene = make_scene(n_voxels=128 ** 3)
        model = model_from_scene(scene, dev)
        model.mlp_mode = "tc_f16"
        opt = TableAdam(model, 0.02, 0.001, 0.1)
        used = opt.enable_peer_exchange() if exchange == "peer" else False
        n = 4096
        rays = make_rays(n, 'isotropic', seed=2000 + rank).to(dev)
        target = torch.rand(n, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + rank))
        for _ in range(3):
            for p in model.parameters():
                p.grad = None
            opt.zero_grad()
            rgb = model(rays, is_train=True, seed=1234, ray_index0=rank * n, **RENDER_KW)[0]
            torch.mean((rgb - target) ** 2).backward()
            model.allreduce_gradients(average=True)
            opt.step()
            model.update_coarse_sigma_grid()
        torch.cuda.synchronize()
        tables = model._render_tables().clone()
        mlp = torch.cat([p.detach().flatten() for p in model._param_list()[24:]])
        opt.disable_peer_exchange()
        return used, tables, mlp

    used, t_peer, m_peer = run("peer")
    _, t_nccl, m_nccl = run("nccl")
    res["train_used_peer"] = bool(used)
    res["train_tables_rel_diff"] = maxr(((t_peer - t_nccl).norm() / t_nccl.norm()).item())
    res["train_mlp_rel_diff"] = maxr(((m_peer - m_nccl).norm() / m_nccl.norm()).item())
    chk = t_peer.double().sum().item()
    lo, hi = torch.tensor([chk], device=dev, dtype=torch.float64), torch.tensor([chk], device=dev, dtype=torch.float64)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["train_tables_identical_on_all_ranks"] = bool(lo.item() == hi.item())
    if rank == 0:
        line = json.dumps(res)
        print(line)
        if args.out:
            open(args.out, "w").write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
