#!/usr/bin/env python
"""bench.py — rays/s of the EgoNeRF volume-rendering path on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--mode render|train] [--impl reference]

A "step" = one pass of the hot path (EgoNeRF.forward through the C ABI of libegn_b200; `--mode train` adds the backward
pass and, for N > 1, the NCCL all-reduce of the gradients) over one batch of synthetic rays per GPU.  Rays shard
across ranks with the parameters replicated (weak scaling: every rank renders `rays` rays per step).

  value      device-resident throughput: rays already in HBM, CUDA events around exactly K steps, max over ranks
  e2e        same metric through the public API (`renderer.volume_renderer`) from PINNED HOST rays: H2D of the rays
             and D2H of the result (rgb + depth; loss in train mode) inside the timed region.  `e2e.with_alpha` is the same
             call with `empty_gpu_cache=True`, i.e. what the reference's evaluation does (renderer.py:39-53): rgb, depth AND
             the (N, S) alpha (1 KB per ray) come back to host memory
  roofline   dominant kernel, algorithmic bytes (SURVEY.md §8d tap model) / its CUDA-event duration, vs the measured
             HBM copy bandwidth in MEASURED_PEAKS.json; `issue` = what actually binds the kernel (ncu of the same build)
  train      (default line, every N) BASELINE configs[2] and configs[3] as TRAINING steps at 16 384 rays per GPU:
             forward + backward + gradient exchange (NCCL) + table-space Adam, with the exchange timed by CUDA events
  erp_frame  BASELINE configs[4]: a 256-row tile per GPU of a 4096 x 2048 equirectangular frame, 256 + 512 samples per ray
  cpu_baseline  the CPU oracle (port of the reference's algorithm, oracle/egn_oracle.py) on a bounded sample of the
             same workload on this box's host cores (rank 0, N = 1 only); `--impl reference` times the SAME leg alone

`--impl reference` times that CPU port alone (the reference itself is a Python program that cannot travel to the GPU
box; the oracle is pinned to it by tests/golden).  Nothing here reads /root/reference.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs restated as synthetic inputs
WORKLOADS = {
    "cfg1": dict(desc="synthetic 360 scene, 128^3 Yin-Yang grid [64,72,216], 4096 rays/batch", n_voxels=128 ** 3, rays=4096,
                 scene={}),
    "cfg2": dict(desc="OmniBlender-shape synthetic, 300^3 grid [150,172,516] VM-decomp, 65536 rays/batch", n_voxels=27e6,
                 rays=65536, scene={}),
    "cfg3": dict(desc="Ricoh360-shape synthetic, 300^3 grid + envmap h=1920, 16384 rays/GPU", n_voxels=27e6, rays=16384,
                 scene=dict(near_far=(0.1, 300.), r0=0.05, density_shift=-10., envmap_h=1920)),
    "cfg5": dict(desc="one 256-row tile (1 048 576 rays / GPU) of a 4096x2048 ERP frame, 256 coarse + 512 fine samples/ray, "
                      "rendered in chunks of 65536 rays", n_voxels=27e6, rays=256 * 4096, scene={}, n_coarse=256, n_fine=256,
                 kind="erp", chunk=65536),
}


def algorithmic_bytes_per_ray(S, n_coarse, elem=4, env=False):
    """SURVEY.md §8(d) tap model: every bilinear/linear tap reads C contiguous elements of `elem` bytes
    (1 328 168 B/ray fp32, 664 616 B/ray bf16 at 128 coarse + 256 fine samples)."""
    io = 24 + 16 + 4 * S + (28 + 48 if env else 0)
    return elem * (n_coarse * 288 + S * 1152) + io


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (NVML; nvidia-smi as fallback)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device):
        super().__init__(daemon=True)
        import torch
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = device.index or 0, [], set(), False, None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:        # CUDA_VISIBLE_DEVICES may renumber: resolve through the UUID
                self.h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(device).uuid))
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _sample(self):
        if self.h is not None:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        else:
            import subprocess
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                  "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip().split(",")
            self.samples.append(int(out[0]))
            self.max_mhz = int(out[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                if v.strip() == "Active":
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


CPU_SAMPLE_RAYS = 4096          # rays per CPU-baseline step: the in-line leg and `--impl reference` time the SAME sample


def oracle_rays_per_s(scene, n_rays, repeats, warmup, seed=5, N_COARSE=128, N_FINE=128):
    """CPU port of the reference path (checker code, used here only as the reported CPU baseline)."""
    import torch
    from oracle import egn_oracle as O
    from egonerf_b200.synthetic import make_rays
    torch.set_num_threads(os.cpu_count())
    cfg = O.OracleCfg(aabb=scene.aabb, grid=tuple(scene.grid), r0=scene.r0, near=scene.near_far[0], far=scene.near_far[1],
                      density_shift=scene.density_shift, distance_scale=scene.distance_scale, n_coarse=N_COARSE,
                      n_fine=N_FINE)
    rays = make_rays(n_rays, 'isotropic', seed=seed)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            O.render(scene.state_dict, cfg, rays, False, emission=scene.emission)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return n_rays * len(times) / sum(times), sum(times) / len(times)


def cpu_baseline_record(rps, repeats):
    return {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{CPU_SAMPLE_RAYS} rays of the same scene per step x {repeats} steps, eval forward, torch CPU fp32 "
                      f"({os.cpu_count()} threads); identical leg in-line and under --impl reference"}


FACT_SOURCES = ("egn_fused.cu", "egn_tc.cuh", "egn_shared.cuh", "egn_device.cuh")


def source_sha():
    """Hash of the sources the dominant kernel (egn_fused_fine_kernel) is compiled from: profile-derived numbers (ncu dram
    traffic, issue utilisation) are reported only when they were captured from THIS build of that kernel
    (profiles/kernel_facts.json records the hash they belong to)."""
    import hashlib
    h = hashlib.sha256()
    for f in FACT_SOURCES:
        h.update(open(os.path.join(ROOT, "egonerf_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU per step (default: the workload's batch)")
    ap.add_argument("--mode", default="render", choices=["render", "train"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-adam", action="store_true",
                    help="train mode: torch.optim.Adam over the reference's parameter groups (what an unchanged train.py does) "
                         "instead of the table-space fused Adam (egonerf_b200.optim.TableAdam)")
    ap.add_argument("--no-parity-line", action="store_true", help="skip the extra fp32-parity-mode measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the train / erp_frame sub-records of the default line")
    ap.add_argument("--tables", default="bf16", choices=["f32", "bf16"],
                    help="tables the tcgen05 BACKWARD kernels re-gather from (the throughput forward always reads the half tables)")
    ap.add_argument("--mlp", default="tc_f16", choices=["fp32", "tc_split", "tc_f16", "tc_bf16"],
                    help="arithmetic: exact fp32 FFMA, tcgen05 3-term bf16 split (fp32-equivalent), tc_f16 = throughput mode "
                         "(fused kernel, fp16 operands + fp32 density; tc_bf16 is its old name)")
    ap.add_argument("--grad-dtype", default="f32", choices=["f32", "bf16"], help="dtype of the factor-gradient all-reduce")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1 training: gradient exchange by one kernel over NVLink peer memory (egn_peer_allreduce) or by ncclAllReduce")
    args = ap.parse_args()
    if args.mlp == "tc_bf16":
        args.mlp = "tc_f16"
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    n_rays = args.rays or wl["rays"]
    N_COARSE, N_FINE = wl.get("n_coarse", 128), wl.get("n_fine", 128)
    chunk = min(wl.get("chunk", n_rays), n_rays)
    S = N_COARSE + N_FINE
    config = {"workload": wl["desc"], "rays_per_gpu_per_step": n_rays, "samples_per_ray": f"{N_COARSE} coarse + {S} fine",
              "mode": args.mode, "sharding": f"rays x{world}, grid replicated"}
    metric = "rays/sec (render)" if args.mode == "render" else "rays/sec (train step: fwd + bwd + grad all-reduce + Adam + table refresh)"

    import torch
    from egonerf_b200.synthetic import make_scene, make_rays

    # ---------------------------------------------------------------- reference arm: CPU port on host cores
    if args.impl == "reference":
        if rank != 0:
            return
        scene = make_scene(n_voxels=wl["n_voxels"], **wl["scene"])
        rps, sec = oracle_rays_per_s(scene, CPU_SAMPLE_RAYS, args.steps, args.warmup, N_COARSE=N_COARSE, N_FINE=N_FINE)
        line = {"impl": "reference", "metric": metric, "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cpu_baseline_record(rps, args.steps),
                "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from egonerf_b200 import _lib
    from egonerf_b200.scene_io import model_from_scene, RENDER_KW
    from egonerf_b200.renderer import volume_renderer
    from egonerf_b200.optim import TableAdam
    _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    flush_buf = [None]

    def timed(fn, steps, warmup, clocks=None, flush=False):
        """W warm-up calls, barrier + synchronize, exactly `steps` timed calls under CUDA events, barrier, MAX over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        if clocks is not None:
            clocks.start()
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        else:
            # small working set: evict L2 between timed iterations (untimed 512 MB write), one event pair per step
            if flush_buf[0] is None:
                flush_buf[0] = torch.zeros(128 * 1024 * 1024, device=dev)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in evs:
                flush_buf[0].add_(1.0)
                a.record()
                fn()
                b.record()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        barrier()
        return max_over_ranks(ms)

    def build(workload):
        w = WORKLOADS[workload]
        scene_ = make_scene(n_voxels=w["n_voxels"], **w["scene"])
        model_ = model_from_scene(scene_, dev)
        model_.mlp_mode, model_.table_dtype = args.mlp, args.tables
        return scene_, model_

    # ---- one training step: what train.py:245-357 does per iteration, ray-sharded -------------------------
    class TrainStep:
        def __init__(self, model_, rays_, n, table_adam=True, host_rays=None):
            self.model, self.rays, self.n, self.host_rays = model_, rays_, n, host_rays
            self.target = torch.rand(n, 3, device=dev)
            self.params = [p for p in model_.parameters()] + ([model_.envmap.emission] if model_.envmap is not None else [])
            model_.sparse_env_grad = model_.envmap is not None          # 24 B/ray exchange instead of the dense envmap gradient
            if table_adam:
                self.opt = TableAdam(model_, 0.02, 0.001, 0.1)
                self.opt.grad_allreduce_dtype = torch.bfloat16 if args.grad_dtype == "bf16" else None
                # gradient exchange: one kernel over NVLink peer memory (egn_peer_allreduce) unless --exchange nccl
                self.peer = (world > 1 and args.exchange == "peer" and args.grad_dtype == "f32"
                             and self.opt.enable_peer_exchange())
            else:
                self.opt = torch.optim.Adam(model_.get_optparam_groups(0.02, 0.001, merged=True), betas=(0.9, 0.99), fused=True)
            self.table_adam = table_adam
            self.ar_events = []
            if not table_adam:
                self.peer = False

        def close(self):
            if self.table_adam:
                self.opt.disable_peer_exchange()

        def __call__(self, e2e=False, time_allreduce=False):
            m = self.model
            for p in self.params:
                p.grad = None
            if self.table_adam:
                self.opt.zero_grad()
            kw_ = dict(RENDER_KW)
            if e2e:
                rgb = volume_renderer(self.host_rays, m, chunk=self.n, is_train=True, device=dev, **kw_)[0]
            else:
                rgb = m(self.rays, is_train=True, seed=1234, ray_index0=rank * self.n, **kw_)[0]
            loss = torch.mean((rgb - self.target) ** 2)
            loss.backward()
            if time_allreduce:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            m.allreduce_gradients(average=True)             # world 1: only the local scatter of the sparse envmap gradient
            if time_allreduce:
                b.record()
                self.ar_events.append((a, b))
            self.opt.step()
            m.update_coarse_sigma_grid()
            if e2e:
                loss.item()
            return loss

        def allreduce_ms(self):
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in self.ar_events]
            self.ar_events = []
            return sum(ms) / max(len(ms), 1)

    def train_record(workload, steps):
        """fwd + bwd + gradient exchange + TableAdam at 16 384 rays per GPU; the exchange is timed separately."""
        scene_, model_ = build(workload)
        n = 16384
        rays_ = make_rays(n, 'isotropic', seed=2000 + rank).to(dev)
        st = TrainStep(model_, rays_, n)
        ms = timed(lambda: st(), steps, 3)
        st.ar_events = []
        ms_t = timed(lambda: st(time_allreduce=True), steps, 0)
        ar = max_over_ranks(st.allreduce_ms())
        numel = model_._render_tables().numel() + sum(p.numel() for p in model_._param_list()[24:-1 if model_.envmap is not None else None])
        rec = {"workload": WORKLOADS[workload]["desc"], "value": n * world * steps / (ms * 1e-3), "unit": "rays/s",
               "rays_per_gpu_per_step": n, "ms_per_step": ms / steps, "allreduce_ms": ar,
               "ms_per_step_with_allreduce_events": ms_t / steps,
               "allreduce": {"bytes_dense": int(numel * (2 if args.grad_dtype == "bf16" else 4)), "dtype": args.grad_dtype,
                             "envmap_gradient": ("sparse: all-gather of 24 B/ray + local scatter (egn_envmap_backward)"
                                                 if model_.envmap is not None else None),
                             "what": ("one kernel over NVLink peer memory (egn_peer_allreduce): table-layout factor gradient + the "
                                      "basis / MLP bucket in its tail, reduce-scatter by peer loads, all-gather by peer stores"
                                      if st.peer else
                                      "NCCL all-reduce of the table-layout factor gradient + one flat bucket of basis / MLP gradients")
                                     if world > 1 else "single GPU: no collective (local envmap scatter only)",
                             "path": ("peer" if st.peer else "nccl") if world > 1 else None},
               "optimizer": "TableAdam (egn_adam_tables)", "gpu_launches_per_step": model_.launches_per_train_step(n)}
        st.close()
        del st, model_, scene_
        torch.cuda.empty_cache()
        return rec

    def erp_record(steps):
        """BASELINE configs[4]: every rank renders a 256-row tile (1 048 576 rays) of a 2048 x 4096 equirect frame, rays
        generated on the device from the pose (egn_erp_rays), 256 coarse + 512 fine samples, chunks of 65 536 rays."""
        from egonerf_b200.raybank import erp_rays
        w = WORKLOADS["cfg5"]
        scene_, model_ = build("cfg5")
        rows = 256
        row0 = (rank * rows) % 2048
        c2w = torch.tensor([[1., 0, 0, 0.1], [0, 1., 0, 0.0], [0, 0, 1., -0.2]])      # a pose off the grid centre
        kw_ = dict(RENDER_KW)
        kw_.update(n_coarse=w["n_coarse"], n_fine=w["n_fine"])

        def step():
            rays_ = erp_rays(2048, 4096, c2w, rows=(row0, row0 + rows), device=dev)
            with torch.no_grad():
                for c0 in range(0, rays_.shape[0], w["chunk"]):
                    model_(rays_[c0:c0 + w["chunk"]], is_train=False, ray_index0=c0, **kw_)
        ms = timed(step, steps, 1)
        n = rows * 4096
        rec = {"workload": w["desc"], "value": n * world * steps / (ms * 1e-3), "unit": "rays/s", "rays_per_gpu_per_step": n,
               "ms_per_step": ms / steps, "samples_per_ray": "256 coarse + 512 fine", "frame_rows_per_gpu": rows,
               "full_frame_ms_at_this_rate": (2048 * 4096) / (n * world * steps / (ms * 1e-3)) * 1e3,
               "rays": "generated on the device from the pose (egn_erp_rays): 48 B of pose instead of 24 B/ray over PCIe"}
        del model_, scene_
        torch.cuda.empty_cache()
        return rec

    # ---------------------------------------------------------------- headline workload
    scene, model = build(args.workload)
    config["mlp"] = args.mlp
    config["backward_tables"] = args.tables
    if wl.get("kind") == "erp":          # this rank's row tile of the 2048 x 4096 equirect frame (ray_utils.py:24-40)
        rays_host = make_rays(n_rays, 'erp', erp_hw=(2048, 4096), row0=(rank * (n_rays // 4096)) % 2048).pin_memory()
    else:
        rays_host = make_rays(n_rays, 'isotropic', seed=1000 + rank).pin_memory()
    rays_dev = rays_host.to(dev)
    ray0 = rank * n_rays
    kw = dict(RENDER_KW)
    kw.update(n_coarse=N_COARSE, n_fine=N_FINE)
    train = args.mode == "train"
    if train:
        tstep = TrainStep(model, rays_dev, n_rays, table_adam=not args.torch_adam, host_rays=rays_host)
        config["exchange"] = ("peer-memory kernel (egn_peer_allreduce)" if tstep.peer else "nccl") if world > 1 else None
        config["optimizer"] = ("torch.optim.Adam(fused=True) + egn_unpack_table_grads + egn_pack_tables" if args.torch_adam
                               else "TableAdam (egn_adam_tables: Adam + table refresh in one pass)")

    def step_device():
        if train:
            return tstep()
        with torch.no_grad():
            for c0 in range(0, n_rays, chunk):
                out = model(rays_dev[c0:c0 + chunk], is_train=False, ray_index0=ray0 + c0, **kw)[0]
            return out

    # per-step working set: factor tables + the per-sample workspace one step streams through
    fused = args.mlp == "tc_f16"
    table_mb = (74.0 if fused else 99.0) * (1.0 if wl["n_voxels"] > 1e7 else wl["n_voxels"] / 27e6)
    ws_mb = min(chunk, n_rays) * S * (24 if (fused and not train) else 136) / 1e6 + n_rays * 4 * S / 1e6
    flush = table_mb + ws_mb < 2 * 126
    if flush:
        config["l2"] = f"working set {table_mb + ws_mb:.0f} MB fits L2: L2 flushed (512 MB write) between timed iterations, per-step CUDA events"
    else:
        config["l2"] = (f"inputs exceed L2: {table_mb:.0f} MB factor tables + {ws_mb:.0f} MB of per-sample state / outputs streamed "
                        "every step (126 MB L2); the tables themselves stay L2-resident, as in any render loop")
    out_rgb = torch.empty(n_rays, 3).pin_memory()
    out_depth = torch.empty(n_rays).pin_memory()

    def step_e2e():
        if train:
            return tstep(e2e=True)
        with torch.no_grad():
            rgb, depth, _, _, _ = volume_renderer(rays_host, model, chunk=chunk, is_train=False, device=dev, **kw)
        out_rgb.copy_(rgb, non_blocking=True)
        out_depth.copy_(depth, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def step_e2e_alpha():       # the reference's evaluation call (renderer.py:129-134, empty_gpu_cache=True): everything to the host
        with torch.no_grad():
            return volume_renderer(rays_host, model, chunk=chunk, is_train=False, device=dev, empty_gpu_cache=True, **kw)

    # the volume_renderer mirror prints the reference's "elapsed time per image" line (renderer.py:75): keep stdout clean
    with contextlib.redirect_stdout(io.StringIO()):
        sampler = ClockSampler(dev)
        ms = timed(step_device, args.steps, args.warmup, sampler, flush=flush)
        clocks = sampler.result()
        ms_e2e = timed(step_e2e, args.steps, 2, flush=flush)
        ms_e2e_alpha = None if train else timed(step_e2e_alpha, args.steps, 2, flush=flush)

    total_rays = n_rays * world * args.steps
    value = total_rays / (ms * 1e-3)
    e2e_value = total_rays / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: per-stage CUDA-event times over the same workload (render stages) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    env = scene.emission is not None
    # algorithmic bytes per fine sample (tap model, SURVEY.md 8d: every tap reads its channels): fp32 tables 18 x 64 x 4 B;
    # half tables 18 x (16 density x 4 B + 48 appearance x 2 B)
    fine_tap_bytes = 18 * (16 * 4 + 48 * 2) if fused else 18 * 64 * 4
    stage_names = ["sampler(coarse+cdf+sort)", "fused fine pass: gather+basis+mlp (egn_fused_fine_kernel)" if fused else
                   "fine_gather+basis (egn_gather_kernel)", "mlp_decode", "composite"]
    stage_alg_bytes = [n_rays * (N_COARSE * 288 * 4 + 24 + 4 * S), n_rays * (S * fine_tap_bytes + 4 * S + (12 * S if fused else 0)),
                       n_rays * S * (28 + 3) * 4, n_rays * (S * (4 + 4 + 12) + 16 + 4 * S)]
    stage_ms = model.stage_times(rays_dev[:chunk], repeats=max(3, min(args.steps, 10)), **kw)
    stage_ms = [x * (n_rays / chunk) for x in stage_ms]
    dom = max(range(len(stage_ms)), key=lambda i: stage_ms[i])
    achieved = stage_alg_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9
    b_ray = N_COARSE * 288 * 4 + S * fine_tap_bytes + 24 + 16 + 4 * S + (28 + 48 if env else 0)
    # numbers that only a profiler sees (dram bytes, issue-slot utilisation, instructions per sample): taken from the committed
    # ncu capture IF it was made from this very build of the kernels, else null -- never a stale constant
    facts, traffic, issue = None, None, None
    fpath = os.path.join(ROOT, "profiles", "kernel_facts.json")
    if os.path.isfile(fpath):
        facts = json.load(open(fpath))
        rec = facts.get("kernels", {}).get(f"{args.workload}:{args.mlp}:{stage_names[dom].split(' (')[0].split(':')[0]}")
        if rec and facts.get("source_sha") == source_sha() and rec.get("rays") == n_rays:
            traffic, issue = rec.get("dram_bytes_per_launch"), rec.get("issue")
    roofline = {"bound": "hbm", "kernel": stage_names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "issue": issue,
                "note": "algorithmic (tap-model) bytes; the factor tables (74 MB half tables / 99 MB fp32) are L2-resident, so most "
                        "of these bytes are served by L2, not HBM: the binding resources are the L1 data pipe (LSU wavefronts) and "
                        "instruction issue -- see `issue` (ncu of this build, null if the committed capture is of another build) "
                        "and profiles/",
                "stage_ms": dict(zip(stage_names, [round(x, 4) for x in stage_ms])),
                "whole_path": {"bytes_per_ray": b_ray, "achieved": value / world * b_ray / 1e9,
                               "frac": value / world * b_ray / 1e9 / peak}}

    dtype = "f16" if fused else "f32"
    config["arithmetic"] = {"fp32": "fp32 everywhere (FFMA MLP)",
                            "tc_split": "fp32 tables; tcgen05 MLP and mma.sync basis with 3-term bf16 split (fp32-equivalent), fp32 accumulate",
                            "tc_f16": "density channels, alpha, transmittance and compositing in fp32; appearance tables, packed-half2 "
                                      "interpolation and tcgen05 MMA operands in fp16 with fp32 accumulate (rgb within 1e-4 of the "
                                      "reference, tests/test_gpu_tc.py); backward: tcgen05 kernels, fp16 operands with a launch-wide power-of-two gradient scale"}[args.mlp]
    launches = (model.launches_per_forward(S) * (-(-n_rays // chunk)) if not train else model.launches_per_train_step(n_rays)) * args.steps
    e2e = {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n_rays * 24 * world,
           "d2h_bytes_per_step": (n_rays * 16 if not train else 4) * world}
    if ms_e2e_alpha is not None:
        acols = S + (1 if env else 0)
        e2e["with_alpha"] = {"value": total_rays / (ms_e2e_alpha * 1e-3), "unit": "rays/s",
                             "d2h_bytes_per_step": n_rays * (16 + 4 * acols + (24 if env else 0)) * world,
                             "what": "volume_renderer(..., empty_gpu_cache=True) as renderer.evaluation calls it (renderer.py:39-53,129-134): "
                                     "rgb, depth and the (N, S) alpha return to pinned host memory, chunk c's copy overlapping chunk c+1"}
    line = {"metric": metric, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline}
    # the same workload in the fp32-parity mode (fp32 tables, tensor-core MLP with the 3-term split), reported alongside
    if fused and not train and not args.no_parity_line:
        model.mlp_mode, model.table_dtype = "tc_split", "f32"
        with contextlib.redirect_stdout(io.StringIO()):
            ms_p = timed(step_device, max(3, args.steps // 2), 3, flush=flush)
        st_p = [x * (n_rays / chunk) for x in model.stage_times(rays_dev[:chunk], repeats=3, **kw)]
        line["parity_mode"] = {"value": n_rays * world * max(3, args.steps // 2) / (ms_p * 1e-3), "unit": "rays/s",
                               "dtype": "f32", "arithmetic": "fp32 tables, tcgen05 MLP + mma.sync basis with 3-term bf16 split (rgb within 1e-4 of the reference)",
                               "stage_ms": dict(zip(["sampler", "gather+basis", "mlp", "composite"], [round(x, 4) for x in st_p]))}
        model.mlp_mode, model.table_dtype = args.mlp, args.tables
    # BASELINE configs[2], [3] as training steps and configs[4] (ERP frame tiles), at every N: the driver's scaling run then
    # records the step that contains the collective, not only the collective-free render
    if not train and not args.no_extras and args.workload == "cfg2":
        del model
        torch.cuda.empty_cache()
        k_steps = max(3, min(args.steps, 10))
        with contextlib.redirect_stdout(io.StringIO()):
            line["train"] = {"cfg2": train_record("cfg2", k_steps), "cfg3": train_record("cfg3", k_steps)}
            line["erp_frame"] = erp_record(max(2, min(args.steps, 3)))
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            rps, _ = oracle_rays_per_s(scene, CPU_SAMPLE_RAYS, 3, 1, N_COARSE=N_COARSE, N_FINE=N_FINE)
            line["cpu_baseline"] = cpu_baseline_record(rps, 3)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
