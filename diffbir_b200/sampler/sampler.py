"""Spaced (DDPM-respaced) and DDIM samplers — counterparts of the reference's
diffbir.sampler.{SpacedSampler, DDIMSampler} (sampler/sampler.py, spaced_sampler.py,
ddim_sampler.py) with the same constructors and `sample(...)` signature.

When `model` is a diffbir_b200 ControlLDM the loop runs on the kernel engine:
  * cond and uncond branches (and, when tiled, all latent tiles) are ONE batched forward,
    captured once in a CUDA graph and replayed every step;
  * text K/V and every time embedding are computed before the loop;
  * CFG mix + x0 + posterior/DDIM update is one fused kernel per step (dbir_sampler_step);
  * tiles are gathered / Gaussian-blended on the device in the reference's accumulation order;
    with torch.distributed initialised the tiles are sharded round-robin over the ranks and
    re-assembled by a single all-gather per step (every rank then blends and updates the same
    full latent with the same RNG stream).
Noise is drawn with torch.randn_like once per step exactly like the reference
(spaced_sampler.py:181, ddim_sampler.py:143): the RNG consumption is the reference's. Outputs are
bit-reproducible across runs and processes in tiled mode and with DBIR_DETERMINISTIC=1 (batch-invariant
kernel plans); otherwise the timing-based split-K choice of dbir_gemm may differ between processes
(rounding-level differences).
Any other callable `model(x, t, cond)` takes a plain PyTorch loop with identical arithmetic.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import lib
from ..utils.common import gaussian_weights, sliding_windows


def space_timesteps(num_timesteps: int, steps: int) -> np.ndarray:
    """space_timesteps(num_timesteps, str(steps)) — spaced_sampler.py:14-64 (single section)."""
    stride = 1.0 if steps <= 1 else (num_timesteps - 1) / (steps - 1)
    out, cur = [], 0.0
    for _ in range(steps):
        out.append(round(cur))
        cur += stride
    return np.array(sorted(set(out)), dtype=np.int32)


# ---- tile sharding (multi-GPU tiled sampling; host-side logic shared with the CPU/gloo test) ----
def tiles_of_rank(num_tiles: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: tile t belongs to rank t % world, local slot t // world."""
    return list(range(rank, num_tiles, world))


def tile_slots(num_tiles: int, world: int) -> int:
    return (num_tiles + world - 1) // world


def assemble_gathered(recv: torch.Tensor) -> torch.Tensor:
    """all_gather output [world, nbr, slots, ...] -> [nbr, slots*world, ...] in global tile order
    (tile t = slot*world + rank); entries >= num_tiles are padding."""
    world, nbr, slots = recv.shape[:3]
    rest = recv.shape[3:]
    perm = (1, 2, 0) + tuple(range(3, recv.dim()))
    return recv.permute(*perm).reshape(nbr, slots * world, *rest).contiguous()


def _sync_from_rank0(t: torch.Tensor, dev) -> torch.Tensor:
    import torch.distributed as dist
    t = t.to(dev).contiguous()
    dist.broadcast(t, src=0)
    return t


def _sync_rng_from_rank0(dev) -> None:
    """All ranks continue with rank 0's generator state (per-step randn_like must agree)."""
    import torch.distributed as dist
    dev = torch.device(dev)
    if dev.type == "cuda":
        st = torch.cuda.get_rng_state(dev).to(dev)
        dist.broadcast(st, src=0)
        torch.cuda.set_rng_state(st.cpu(), dev)
    else:
        st = torch.get_rng_state()
        dist.broadcast(st, src=0)
        torch.set_rng_state(st)


class Sampler:
    time_collective = False        # record CUDA events around the per-step all-gather (bench.py sets it)

    def __init__(self, betas: np.ndarray, parameterization: str, rescale_cfg: bool):
        self.num_timesteps = len(betas)
        self.training_betas = betas
        self.training_alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.parameterization = parameterization
        self.rescale_cfg = rescale_cfg
        self.shard_tiles = True     # tiled sampling: shard tiles over torch.distributed ranks
        self.last_stats: dict = {}

    def get_cfg_scale(self, default_cfg_scale: float, model_t: int) -> float:
        """Cosine CFG ramp — sampler/sampler.py:31-38."""
        if self.rescale_cfg and default_cfg_scale > 1:
            return 1 + default_cfg_scale * ((1 - math.cos(math.pi * ((1000 - model_t) / 1000) ** 5.0)) / 2)
        return default_cfg_scale

    # ---- subclass interface -------------------------------------------------------------
    def make_schedule(self, steps: int):
        raise NotImplementedError

    def _step_coefs(self) -> Tuple[np.ndarray, torch.Tensor, int]:
        """(model timesteps ascending, fp32 coefficient rows [S, 8] per table index, kernel mode)"""
        raise NotImplementedError

    # ---- shared driver --------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, model, device, steps: int, x_size: Tuple[int], cond: Dict[str, torch.Tensor],
               uncond: Optional[Dict[str, torch.Tensor]], cfg_scale: float, tiled: bool = False,
               tile_size: int = -1, tile_stride: int = -1, x_T: Optional[torch.Tensor] = None,
               progress: bool = True) -> torch.Tensor:
        self.make_schedule(steps)
        ts, coefs, mode = self._step_coefs()
        if x_T is None:
            x_T = torch.randn(x_size, device=device, dtype=torch.float32)
        from ..model.cldm import ControlLDM
        if isinstance(model, ControlLDM):
            return self._sample_engine(model, ts, coefs, mode, x_T, cond, uncond, cfg_scale, tiled,
                                       tile_size, tile_stride)
        return self._sample_generic(model, ts, coefs, mode, x_T, cond, uncond, cfg_scale, tiled,
                                    tile_size, tile_stride)

    # ---- kernel-engine path -----------------------------------------------------------------
    def _sample_engine(self, model, ts, coefs, mode, x_T, cond, uncond, cfg_scale, tiled, tile_size,
                       tile_stride):
        model._build()
        eng = model.engine
        dev = eng.dev
        x = x_T.to(dev, torch.float32).contiguous().clone()
        B, C, H, W = x.shape
        use_cfg = not (uncond is None or cfg_scale == 1.0)
        conds = [cond, uncond] if use_cfg else [cond]
        nbr = len(conds)
        coefs = coefs.to(dev)
        order = list(range(len(ts)))[::-1]                     # table index of each loop iteration
        model_ts = [int(ts[i]) for i in order]

        import torch.distributed as dist
        world, rank = ((dist.get_world_size(), dist.get_rank())
                       if (tiled and self.shard_tiles and dist.is_available() and dist.is_initialized()) else (1, 0))
        if world > 1:
            # Every rank must hold the same latent, condition and noise stream: rank 0's are authoritative
            # (ranks seeded differently would otherwise blend tiles of diverging latents silently).
            x = _sync_from_rank0(x, dev)
            conds = [dict(cd, c_img=_sync_from_rank0(cd["c_img"].to(dev, torch.float32).contiguous(), dev))
                     for cd in conds]
            _sync_rng_from_rank0(dev)

        if tiled:
            wins = sliding_windows(H, W, tile_size, tile_stride)
            T = len(wins)
            all_coords = torch.tensor([[a, c] for a, _, c, _ in wins], dtype=torch.int32, device=dev)
            mine = tiles_of_rank(T, rank, world)               # round-robin tile ownership (may be empty: T < world)
            slots = tile_slots(T, world)
            Tl, ts_ = len(mine), tile_size
            my_coords = all_coords[mine].contiguous() if Tl else all_coords[:0]
            wts = torch.tensor(gaussian_weights(ts_, ts_), dtype=torch.float32, device=dev)
            nb = nbr * Tl * B
            c_img = torch.empty(nbr, Tl * B, C, ts_, ts_, device=dev)
            for j, cd in enumerate(conds):
                if Tl:
                    lib.tile_gather(cd["c_img"].to(dev, torch.float32).contiguous(), B, C, H, W, my_coords,
                                    Tl, ts_, c_img[j])
            c_img = c_img.view(nb, C, ts_, ts_)
            ctx = torch.cat([cd["c_txt"].to(dev, torch.float32).repeat(Tl, 1, 1) for cd in conds], 0)
            gh = gw = ts_
            send = torch.zeros(nbr, slots, B, C, ts_, ts_, device=dev)
            recv = torch.empty(world, nbr, slots, B, C, ts_, ts_, device=dev) if world > 1 else None
            eps_full = torch.empty(nbr, B, C, H, W, device=dev)
        else:
            nb = nbr * B
            c_img = torch.cat([cd["c_img"].to(dev, torch.float32) for cd in conds], 0).contiguous()
            ctx = torch.cat([cd["c_txt"].to(dev, torch.float32) for cd in conds], 0)
            gh, gw = H, W
        scales = [float(s) for s in model.control_scales]
        # Tiled sampling pins the batch-invariant kernel plans (no split-K, whole attention tiles per
        # CTA): a tile's eps then has the same bits whatever the per-rank batch is, so the sharded run
        # is bit-identical to the single-rank run.
        eng.batch_invariant = bool(tiled) or eng.deterministic
        if nb > 0:
            eng.set_context(ctx)
            eng.set_timesteps(model_ts, nb)
            model._ctx_ref = model._t_key = None                # generic-path caches are now stale
            eng.load_step(0)
            graph, x_in, c_img_buf, eps, graph_launches = eng.graphed_forward(nb, C, gh, gw, scales)
            c_img_buf.copy_(c_img)

        def fill_inputs():
            if tiled:
                v = x_in.view(nbr, Tl * B, C, ts_, ts_)
                lib.tile_gather(x, B, C, H, W, my_coords, Tl, ts_, v[0])
                for j in range(1, nbr):
                    v[j].copy_(v[0])
            else:
                v = x_in.view(nbr, B, C, H, W)
                for j in range(nbr):
                    v[j].copy_(x)

        x_next = torch.empty_like(x)
        self.last_stats = dict(world=world, tiles=(T if tiled else 0), tiles_this_rank=(Tl if tiled else 0),
                               forwards_per_step=nb)
        for it, tab_idx in enumerate(order):
            if nb > 0:
                eng.load_step(it)
                fill_inputs()
                graph.replay()
                lib.count_launch(graph_launches)
            if tiled:
                if world > 1:
                    if Tl:
                        send[:, :Tl].copy_(eps.view(nbr, Tl, B, C, ts_, ts_))
                    if self.time_collective:
                        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        ev0.record()
                    dist.all_gather_into_tensor(recv, send)
                    if self.time_collective:
                        ev1.record()
                        self.last_stats.setdefault("allgather_events", []).append((ev0, ev1))
                    tiles = assemble_gathered(recv)             # global tile order, padding at the end
                else:
                    tiles = eps.view(nbr, Tl, B, C, ts_, ts_)
                for j in range(nbr):
                    lib.tile_blend(tiles[j], B, C, H, W, all_coords, T, ts_, wts, eps_full[j])
                e_c, e_u = eps_full[0], (eps_full[1] if use_cfg else None)
            else:
                ev = eps.view(nbr, B, C, H, W)
                e_c, e_u = ev[0], (ev[1] if use_cfg else None)
            noise = torch.randn_like(x)
            cur_cfg = self.get_cfg_scale(cfg_scale, model_ts[it])
            if cur_cfg == 1.0:
                e_u = None      # the reference takes the cond-only branch on such steps (spaced_sampler.py:150-152)
            lib.sampler_step(e_c, e_u, cur_cfg, x, noise, coefs[tab_idx], mode, x.numel(), x_next)
            x, x_next = x_next, x
        return x

    # ---- plain PyTorch path for foreign models --------------------------------------------
    def _sample_generic(self, model, ts, coefs, mode, x_T, cond, uncond, cfg_scale, tiled, tile_size,
                        tile_stride):
        x = x_T
        dev = x.device
        coefs = coefs.to(dev)
        fwd = model
        if tiled:
            def fwd(xx, tt, cd):
                out = torch.zeros_like(xx)
                cnt = torch.zeros_like(xx)
                w = torch.tensor(gaussian_weights(tile_size, tile_size)[None, None], dtype=xx.dtype, device=dev)
                for a, b, c, d in sliding_windows(xx.shape[2], xx.shape[3], tile_size, tile_stride):
                    out[..., a:b, c:d] += model(xx[..., a:b, c:d], tt, {"c_txt": cd["c_txt"], "c_img": cd["c_img"][..., a:b, c:d]}) * w
                    cnt[..., a:b, c:d] += w
                return out / cnt
        for tab_idx in list(range(len(ts)))[::-1]:
            step = int(ts[tab_idx])
            model_t = torch.full((x.shape[0],), step, device=dev, dtype=torch.long)
            s = self.get_cfg_scale(cfg_scale, step)
            if uncond is None or s == 1.0:
                e = fwd(x, model_t, cond)
            else:
                ec, eu = fwd(x, model_t, cond), fwd(x, model_t, uncond)
                e = eu + s * (ec - eu)
            c = coefs[tab_idx]
            noise = torch.randn_like(x)
            if mode in (0, 1):
                x0 = c[0] * x - c[1] * e
                x = (c[2] * x0 + c[3] * x) + c[4] * noise
            else:
                if mode == 3:
                    e = c[5] * e + c[1] * x
                x0 = (x - c[1] * e) / c[0]
                x = c[2] * x0 + c[3] * e + c[4] * noise
        return x


class SpacedSampler(Sampler):
    def make_schedule(self, num_steps: int) -> None:
        """spaced_sampler.py:77-116 — respaced betas and posterior tables in fp64."""
        used = space_timesteps(self.num_timesteps, num_steps)
        betas, last = [], 1.0
        for i in used:
            betas.append(1 - self.training_alphas_cumprod[i] / last)
            last = self.training_alphas_cumprod[i]
        self.timesteps = used
        b = np.array(betas, dtype=np.float64)
        a = 1.0 - b
        ac = np.cumprod(a, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        with np.errstate(divide="ignore"):
            self.tables = dict(
                sqrt_alphas_cumprod=np.sqrt(ac), sqrt_one_minus_alphas_cumprod=np.sqrt(1 - ac),
                sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac), sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
                posterior_variance=b * (1.0 - ac_prev) / (1.0 - ac),
                posterior_mean_coef1=b * np.sqrt(ac_prev) / (1.0 - ac),
                posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(a) / (1.0 - ac))

    def _step_coefs(self):
        tb = {k: torch.tensor(v, dtype=torch.float32) for k, v in self.tables.items()}
        S = len(self.timesteps)
        c = torch.zeros(S, 8, dtype=torch.float32)
        if self.parameterization == "eps":
            c[:, 0], c[:, 1], mode = tb["sqrt_recip_alphas_cumprod"], tb["sqrt_recipm1_alphas_cumprod"], 0
        else:
            c[:, 0], c[:, 1], mode = tb["sqrt_alphas_cumprod"], tb["sqrt_one_minus_alphas_cumprod"], 1
        c[:, 2], c[:, 3] = tb["posterior_mean_coef1"], tb["posterior_mean_coef2"]
        nonzero = (torch.arange(S) != 0).float()                  # spaced_sampler.py:182
        c[:, 4] = nonzero * torch.sqrt(tb["posterior_variance"])
        return self.timesteps, c, mode


class DDIMSampler(Sampler):
    def __init__(self, betas, parameterization, rescale_cfg, eta: float):
        super().__init__(betas, parameterization, rescale_cfg)
        self.eta = eta

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform"):
        """ddim_sampler.py:13-58, 73-96 (uniform stride, +1 offset)."""
        c = self.num_timesteps // ddim_num_steps
        self.ddim_timesteps = np.asarray(list(range(0, self.num_timesteps, c))) + 1
        ac = self.training_alphas_cumprod
        alphas = ac[self.ddim_timesteps]
        alphas_prev = np.asarray([ac[0]] + ac[self.ddim_timesteps[:-1]].tolist())
        sigmas = self.eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
        self.tables = dict(ddim_sigmas=sigmas, ddim_alphas=alphas, ddim_alphas_prev=alphas_prev,
                           ddim_sqrt_alphas=np.sqrt(alphas), ddim_sqrt_one_minus_alphas=np.sqrt(1.0 - alphas))

    def _step_coefs(self):
        tb = {k: torch.tensor(v, dtype=torch.float32) for k, v in self.tables.items()}
        S = len(self.ddim_timesteps)
        c = torch.zeros(S, 8, dtype=torch.float32)
        c[:, 0] = tb["ddim_alphas"].sqrt()                                       # a_t.sqrt()
        c[:, 1] = tb["ddim_sqrt_one_minus_alphas"]
        c[:, 2] = tb["ddim_alphas_prev"].sqrt()
        c[:, 3] = (1.0 - tb["ddim_alphas_prev"] - tb["ddim_sigmas"] ** 2).sqrt()
        c[:, 4] = tb["ddim_sigmas"]
        c[:, 5] = tb["ddim_sqrt_alphas"]
        return self.ddim_timesteps, c, (2 if self.parameterization == "eps" else 3)
