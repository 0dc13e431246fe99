"""Environment map mirror (reference: models/envmap.py:6-37): equirect (3, 2h, h) emission leaf tensor, bilinear
lookup + sigmoid, executed by libegn_b200 (`egn_envmap_radiance` / `egn_envmap_backward`)."""
import torch

from .. import _lib


class _EnvRadiance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emission, dirs, h):
        lib = _lib.load()
        cfg = _lib.EgnConfig()
        cfg.env_h = h
        out = torch.empty(dirs.shape[0], 3, device=dirs.device)
        _lib.check(lib.egn_envmap_radiance(cfg, emission.data_ptr(), dirs.data_ptr(), dirs.shape[0], out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(emission, dirs)
        ctx.h = h
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        emission, dirs = ctx.saved_tensors
        cfg = _lib.EgnConfig()
        cfg.env_h = ctx.h
        d_em = torch.zeros_like(emission)
        d_out = d_out.contiguous().float()
        _lib.check(lib.egn_envmap_backward(cfg, emission.data_ptr(), dirs.data_ptr(), dirs.shape[0], d_out.data_ptr(),
                                           d_em.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return d_em, None, None


class EnvironmentMap:
    def __init__(self, h=1000, init_strategy="random", device="cuda"):
        if init_strategy == "random":
            self.emission = torch.rand((3, 2 * h, h), requires_grad=True, device=device)
        elif init_strategy == "zero":
            self.emission = torch.zeros((3, 2 * h, h), requires_grad=True, device=device)
        else:
            raise ValueError("Unknown environment map initialization: {}".format(init_strategy))

    def get_radiance(self, direction):
        """(N,3) directions -> (N,3) radiance in (0,1)  (models/envmap.py:25-34)."""
        if not direction.is_cuda:
            raise RuntimeError("egonerf_b200: envmap lookup needs CUDA tensors — there is no CPU fallback")
        d = direction.detach().contiguous().float()
        return _EnvRadiance.apply(self.emission, d, self.emission.shape[2])

    def load_envmap(self, emission, device):
        self.emission = torch.as_tensor(emission).detach().clone().to(device).requires_grad_(True)
