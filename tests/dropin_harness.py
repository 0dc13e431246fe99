"""Runs the UNMODIFIED reference `train.py` with `shim/` ahead of the reference checkout on sys.path.

TEST INFRASTRUCTURE.  Executed as a subprocess by tests/test_shim_cpu.py (needs /root/reference, so never on the GPU box):

    python tests/dropin_harness.py <reference_root> <workdir>

What it does, in the order `python train.py --config ...` would:
  1. stubs the third-party packages the reference hard-imports and this image lacks (none is used by the path:
     matplotlib, kornia, imageio, plyfile, skimage, lpips) and provides a minimal `configargparse` (argparse + the
     `key = value` / `[a, b]` config-file syntax) so that the reference's own opt.py builds `args` from its own config chain
     configs/EgoNeRF/omniblender/barbershop/default.txt -> common.txt -> ../common_indoor.txt -> ../common.txt;
  2. executes train.py lines 1-20 (the import block) verbatim, then the whole module;
  3. registers a small synthetic dataset under `dataset_dict['omniblender']` (no dataset is reachable offline);
  4. calls the reference's `train(args)`: dataset -> coordinates -> `EgoNeRF(...)` -> optimiser (train.py:118-186) -> first
     `renderer(...)` call of the loop (train.py:253).  Without a GPU that call must end in the drop-in's own
     "no CPU fallback" error; with a GPU the loop runs `--n_iters` iterations, saves a checkpoint and evaluates.
Prints one line per milestone; the caller asserts on them.
"""
import argparse
import os
import sys
import types


def install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "kornia", "imageio", "plyfile", "skimage", "skimage.measure",
                 "skimage.metrics", "lpips", "configargparse"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia"].create_meshgrid = getattr(sys.modules["kornia"], "create_meshgrid", lambda *a, **k: None)
    for a in ("PlyData", "PlyElement"):
        if not hasattr(sys.modules["plyfile"], a):
            setattr(sys.modules["plyfile"], a, object)
    if not hasattr(sys.modules["skimage.measure"], "marching_cubes"):
        sys.modules["skimage.measure"].marching_cubes = lambda *a, **k: None
    if not hasattr(sys.modules["imageio"], "imwrite"):
        sys.modules["imageio"].imwrite = lambda path, img: open(path, "wb").write(b"stub")
    cap = sys.modules["configargparse"]
    if hasattr(cap, "ArgumentParser"):
        return

    class ArgumentParser(argparse.ArgumentParser):
        """argparse + configargparse's config files: `key = value`, bare `flag`, `key = [a, b]` for action='append';
        precedence command line > --config file > default_config_files."""

        def __init__(self, *a, default_config_files=None, **k):
            super().__init__(*a, **k)
            self._default_files = list(default_config_files or [])
            self._config_dests = []

        def add_argument(self, *a, is_config_file=False, **k):
            act = super().add_argument(*a, **k)
            if is_config_file:
                self._config_dests.append(act.dest)
            return act

        @staticmethod
        def _file_args(path):
            out = []
            for line in open(path):
                line = line.split("#")[0].strip()
                if not line:
                    continue
                if "=" not in line:
                    out.append("--" + line)
                    continue
                key, val = (s.strip() for s in line.split("=", 1))
                if val.startswith("["):
                    for item in val.strip("[]").split(","):
                        if item.strip():
                            out += ["--" + key, item.strip()]
                elif val.lower() in ("true",):
                    out.append("--" + key) if key in ("exp_sampling",) else out.extend(["--" + key, val])
                else:
                    out += ["--" + key, val]
            return out

        def parse_args(self, args=None, namespace=None):
            if isinstance(args, str):
                args = args.split()
            args = list(sys.argv[1:] if args is None else args)
            pre, _ = super().parse_known_args(args)
            files = list(self._default_files) + [getattr(pre, d) for d in self._config_dests if getattr(pre, d, None)]
            cli_keys = {a for a in args if a.startswith("--")}     # the command line REPLACES a file's values
            # later files / the --config file override earlier defaults for scalar options by argparse's last-wins rule;
            # for append-type options keep only the LAST file that sets them
            seen_append = {}
            for act in self._actions:
                if isinstance(act, argparse._AppendAction):
                    for opt in act.option_strings:
                        seen_append[opt] = None
            per_file = [self._file_args(f) for f in files]
            last_file_with = {}
            for fi, fa in enumerate(per_file):
                for tok in fa:
                    if tok in seen_append:
                        last_file_with[tok] = fi
            final = []
            for fi, fa in enumerate(per_file):
                i = 0
                while i < len(fa):
                    tok = fa[i]
                    has_val = i + 1 < len(fa) and not fa[i + 1].startswith("--")
                    drop = tok in cli_keys or (tok in seen_append and last_file_with[tok] != fi)
                    if not drop:
                        final.append(tok)
                        if has_val:
                            final.append(fa[i + 1])
                    i += 2 if has_val else 1
            return super().parse_args(final + args, namespace)

    cap.ArgumentParser = ArgumentParser


def make_dataset_class():
    import torch
    from egonerf_b200.synthetic import make_rays

    class SyntheticOmniDataset:
        """Attribute surface of dataLoader/dataset_omniblender.py that train.py / renderer.evaluation touch."""

        def __init__(self, data_dir, split='train', downsample=1.0, is_stack=False, use_gt_depth=False, near_far=None,
                     localization_method='colmap', skip=1, **kw):
            self.near_far = list(near_far) if near_far else [0.01, 15.0]
            self.white_bg = False
            self.img_wh = (32, 16)
            self.roi = [0., 1., 0., 1.]
            half = 0.5 + self.near_far[1]                                    # dataset_omniblender.py:24-32
            self.scene_bbox = torch.tensor([[-half] * 3, [half] * 3])
            n_img = 2
            n = n_img * self.img_wh[0] * self.img_wh[1]
            g = torch.Generator().manual_seed(5 if split == 'train' else 6)
            rays = make_rays(n, 'isotropic', seed=11 if split == 'train' else 12)
            rgbs = torch.rand(n, 3, generator=g)
            if is_stack:
                self.all_rays = rays.view(n_img, -1, 6)
                self.all_rgbs = rgbs.view(n_img, self.img_wh[1], self.img_wh[0], 3)
            else:
                self.all_rays, self.all_rgbs = rays, rgbs

    return SyntheticOmniDataset


def main():
    ref_root, workdir = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2])
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.dont_write_bytecode = True                     # the reference mount is read-only
    install_stubs()
    os.chdir(ref_root)                                 # `python train.py` runs from the checkout
    sys.path.insert(0, ref_root)                       # what `python train.py` puts first ...
    sys.path.insert(0, os.path.join(repo, "shim"))     # ... and PYTHONPATH=<repo>/shim puts ahead of it
    src = open(os.path.join(ref_root, "train.py")).read()
    ns = {"__name__": "train_import_block"}
    exec(compile("\n".join(src.split("\n")[:20]), "train.py", "exec"), ns)         # train.py:1-20 verbatim
    print("IMPORT_BLOCK_OK", ns["volume_renderer"].__module__, ns["evaluation"].__module__, ns["EgoNeRF"].__module__,
          ns["TensorVMSplit"].__module__, ns["coordinates_dict"]["yinyang"].__module__)
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_train", os.path.join(ref_root, "train.py"))
    train_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(train_mod)                                             # the whole file, unmodified
    train_mod.dataset_dict["omniblender"] = make_dataset_class()
    import torch
    torch.manual_seed(20221028)                                                    # train.py:412
    n_iters = os.environ.get("EGN_DROPIN_ITERS", "3")
    sys.argv = ["train.py", "--config", os.path.join(ref_root, "configs/EgoNeRF/omniblender/barbershop/default.txt"),
                "--basedir", workdir, "--expname", "dropin", "--datadir", workdir, "--n_iters", n_iters,
                "--batch_size", "256", "--N_voxel_init", str(40 ** 3), "--N_voxel_final", str(40 ** 3),
                "--progress_refresh_rate", "1", "--vis_list", "2", "--N_vis", "1", "--i_weights", "2"]
    os.makedirs(os.path.join(workdir, "dropin"), exist_ok=True)
    args = train_mod.recursive_config_parser().parse_args()
    print("ARGS_OK", args.model_name, args.coordinates_name, args.n_lamb_sigma, args.n_lamb_sh, args.near_far, args.r0,
          args.density_shift, args.interval_th, args.resampling, args.n_coarse, args.n_fine, args.shadingMode, args.view_pe)
    # spy on the construction milestones without touching the reference code
    from egonerf_b200.models import EgoNeRF as egn_mod
    orig_init = egn_mod.EgoNeRF.__init__

    def spy_init(self, *a, **k):
        orig_init(self, *a, **k)
        print("MODEL_OK", type(self).__module__, self.gridSize.tolist(), sum(p.numel() for p in self.parameters()),
              len(self.get_optparam_groups()))
    egn_mod.EgoNeRF.__init__ = spy_init
    try:
        train_mod.train(args)
        print("TRAIN_DONE", sorted(f for f in os.listdir(os.path.join(workdir, "dropin")) if f.endswith(".th")))
    except RuntimeError as e:
        print("TRAIN_STOPPED_AT", type(e).__name__, str(e)[:160])


if __name__ == "__main__":
    main()
