#!/usr/bin/env python
"""End-to-end seconds per stylised image (BASELINE.json metric part ii, configs[1]): the reference driver loop
(run_strotss.py:43-161) restated in torch AROUND the B200 loss path, on synthetic images of the shapes the
shipped samples produce (content 321x481 -> 341x512 at the last scale, style 1600x1200 -> 512x384).

What is measured and what is not:
  * loss path (sampler, loss + gradient): this repo's CUDA kernels -- the thing being built;
  * VGG16 forward/backward: torch/cuDNN with RANDOM weights (the reference's weights are fetched from a URL and are
    not available offline) -- a stand-in for "the reference's TF/cuDNN path", timing only;
  * pyramid fold (+ backward) and RMSprop on the six Laplacian-pyramid variables: this repo's kernels in
    the default graph mode, torch ops with --eager.
The stylised image is therefore meaningless; only the time is reported.  One JSON line on stdout.

    python bench_e2e.py [--max_iter 200] [--level 4] [--sample 1024] [--pairs 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench_e2e.py --gpus 8 --level 5 --pairs 64
        (BASELINE configs[4]: 64 independent pairs at 1024 px, one process per GPU as the reference's --gpu_id implies)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import strotss_tensorflow_b200 as S  # noqa: E402

# (out_channels, taps_this_layer?) of VGG16 up to block5_conv3; 'M' = 2x2 max-pool (nn/model.py:7-15 lists the taps)
_VGG16 = [(64, True), (64, True), "M", (128, True), (128, True), "M", (256, True), (256, True), (256, True), "M",
          (512, False), (512, False), (512, True), "M", (512, False), (512, False), (512, True)]


class VGG16Features(torch.nn.Module):
    """Conv stack of Keras VGG16(include_top=False) returning the nine post-ReLU maps the reference taps."""

    def __init__(self):
        super().__init__()
        convs, cin = [], 3
        for item in _VGG16:
            if item == "M":
                continue
            conv = torch.nn.Conv2d(cin, item[0], 3, padding=1)
            torch.nn.init.kaiming_normal_(conv.weight, nonlinearity="relu")
            torch.nn.init.zeros_(conv.bias)
            convs.append(conv)
            cin = item[0]
        self.convs = torch.nn.ModuleList(convs)
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x):                       # x: (1, 3, H, W) in [0, 1]
        x = (x - self.mean) / self.std
        outs, k = [], 0
        for item in _VGG16:
            if item == "M":
                x = F.max_pool2d(x, 2, 2)
                continue
            x = F.relu(self.convs[k](x))
            k += 1
            if item[1]:
                outs.append(x)
        return outs


def nhwc(x):                                    # (1, C, H, W) -> contiguous (1, H, W, C) for the sampler
    return x.permute(0, 2, 3, 1).contiguous()


def resize_long(img, long_side):                # nn/utils.py:32-37
    h, w = img.shape[-2:]
    f = max(h / long_side, w / long_side)
    return F.interpolate(img, size=(int(h / f), int(w / f)), mode="bilinear", align_corners=False)


def make_laplacian(x, return_down=False):       # nn/strotss_utils.py:139-146
    h, w = x.shape[-2:]
    hd, wd = max(h // 2, 1), max(w // 2, 1)
    down = F.interpolate(x, size=(hd, wd), mode="bilinear", align_corners=False)
    pyr = x - F.interpolate(down, size=(h, w), mode="bilinear", align_corners=False)
    return (pyr, down) if return_down else pyr


def make_pyramid(x, levels=5):                  # nn/strotss_utils.py:149-156
    out, cur = [], x
    for _ in range(levels):
        pyr, cur = make_laplacian(cur, True)
        out.append(pyr)
    out.append(cur)
    return out


def fold_pyramid_torch(xs):                     # nn/strotss_utils.py:159-163 (torch ops, --eager)
    ret = xs[-1]
    for x in reversed(xs[:-1]):
        ret = x + F.interpolate(ret, size=x.shape[-2:], mode="bilinear", align_corners=False)
    return ret


class _Cfg:
    def __init__(self, level, sample, lr=2e-3, eager=False):
        self.level, self.sample, self.lr, self.eager = level, sample, lr, eager


_VGG = {}


def _setup(dev, sample):
    torch.manual_seed(0)
    gen = torch.Generator(device=dev).manual_seed(0)
    content = torch.rand(1, 3, 321, 481, generator=gen, device=dev)
    style = torch.rand(1, 3, 1600, 1200, generator=gen, device=dev)
    vgg = _VGG.get(dev.index)
    if vgg is None:
        vgg = VGG16Features().to(dev).to(memory_format=torch.channels_last)
        _VGG[dev.index] = vgg
    sampling = S.Sampling(sample, torch.Generator().manual_seed(0))

    def feats(img):
        return [nhwc(img)] + [nhwc(f) for f in vgg(img.contiguous(memory_format=torch.channels_last))]

    return content, style, vgg, sampling, feats


def reset():
    """Drop the captured per-scale graphs (and their static buffers)."""
    _SCALE_GRAPHS.clear()
    torch.cuda.empty_cache()


def run_images(dev, level=4, max_iter=200, sample=1024, images=1, warm_iters=3, eager=False, lr=2e-3):
    """`images` stylised images one after the other on `dev` (the job of ONE GPU in BASELINE configs[4]'s throughput mode,
    run_strotss.py:70-71,145-152,176-179: `--gpu_id` = one process per GPU).  An untimed pass of `warm_iters` iterations
    per scale captures the per-scale graphs; the timed images reuse them.  -> dict with seconds_per_image."""
    cfg = _Cfg(level, sample, lr, eager)
    content, style, vgg, sampling, feats = _setup(dev, sample)
    with torch.cuda.device(dev):
        _run(cfg, content, style, sampling, feats, vgg, warm_iters)
        torch.cuda.synchronize(dev)
        t_all = time.perf_counter()
        per_scale = None
        for _ in range(images):
            per_scale = _run(cfg, content, style, sampling, feats, vgg, max_iter)
        torch.cuda.synchronize(dev)
        total = time.perf_counter() - t_all
    return {"seconds": total, "images": images, "seconds_per_image": total / images, "per_scale_last_image": per_scale}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max_iter", type=int, default=200)
    ap.add_argument("--level", type=int, default=4, help="number of scales; 4 = 512 px, 5 = 1024 px long side (run_strotss.py:70-71)")
    ap.add_argument("--sample", type=int, default=1024)
    ap.add_argument("--lr", type=float, default=2e-3)
    ap.add_argument("--pairs", type=int, default=1, help="content/style pairs of the whole job (BASELINE configs[4]: 64), split over the GPUs")
    ap.add_argument("--gpus", type=int, default=1, help="informational; under torchrun every rank drives the GPU LOCAL_RANK")
    ap.add_argument("--eager", action="store_true", help="launch every iteration op by op instead of replaying a CUDA graph")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    mine = args.pairs // world + (1 if rank < args.pairs % world else 0)       # independent jobs: no data-path collective
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_job = time.perf_counter()
    res = run_images(dev, args.level, args.max_iter, args.sample, max(mine, 1), 3, args.eager, args.lr) if mine > 0 else None
    own = res["seconds"] if res else 0.0
    wall = own
    per_rank = [own]
    if world > 1:
        t = torch.tensor([own], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [float(a.item()) for a in allt]
        wall = max(per_rank)
    if rank == 0:
        px = 2 << (4 + args.level)
        print(json.dumps({
            "metric": f"end-to-end seconds per stylised image, {px} px long side, default settings",
            "value": wall / max(1, -(-args.pairs // world)), "unit": "s/image (per GPU; slowest rank)",
            "images_per_s": args.pairs / wall, "pairs": args.pairs, "n_gpus": world, "seconds_timed_per_rank": per_rank,
            "wall_seconds_incl_warmup_and_capture": time.perf_counter() - t_job,
            "higher_is_better": False, "scaling": "weak (independent jobs, one process per GPU, no collective)",
            "data": "synthetic images (content 321x481, style 1600x1200), random VGG16 weights",
            "config": {"level": args.level, "max_iter": args.max_iter, "sample_size": args.sample, "optimizer": "RMSprop(0.99, 1e-8)",
                       "vgg": "torch/cuDNN conv stack, fp32 tensors, channels_last (stand-in for the reference's TF/cuDNN path)",
                       "loss_path": "strotss_tensorflow_b200 (fused sampler + loss/grad kernels)",
                       "launch": "eager (op by op)" if args.eager else
                                 "one CUDA graph per scale: fold + VGG fwd + sampler + loss/grad + VGG bwd + RMSprop, captured once per "
                                 "process (in the warm-up pass) and reused for every image: a new image only copies its content "
                                 "features / initial pyramid into the graph's buffers and re-prepares the style target; sample indices "
                                 "are drawn on the host and copied into the graph's index buffer each iteration; the loss scalar is "
                                 "read back every iteration",
                       "pixel_side": "torch ops" if args.eager else "strotss_pyramid_fold / _fold_backward / strotss_rmsprop_step (this repo)",
                       "warmup": "one untimed pass of 3 iterations per scale (it also captures the per-scale graphs: steady-state "
                                 "per-image time, as for the 2nd..64th image of BASELINE configs[4])"},
            "per_scale": res["per_scale_last_image"] if res else None}))
    if world > 1:
        dist.destroy_process_group()


class ScaleGraph:
    """Everything one scale of the driver loop needs, with static buffers and ONE captured CUDA graph of an iteration
    (run_strotss.py:131-148): fold -> VGG forward -> sampler -> loss + gradient -> VGG backward -> RMSprop."""

    def __init__(self, sampling, vgg_feats, content_feat, style_samples, stylized_nhwc, alpha, lr):
        self.sampling = sampling
        self.content_feat = [t.clone() for t in content_feat]
        self.variables = [v.clone().requires_grad_(True) for v in S.make_laplacian_pyramid(stylized_nhwc, 5)]
        self.opt = S.RMSprop(rho=0.99, epsilon=1e-8, learning_rate=lr)            # run_strotss.py:63
        self.opt.build(self.variables)
        self.loss_fn = S.StrotssLoss(style_samples, alpha)
        self.static_idx = sampling._make_indices(self.content_feat[0], True)
        self.base_cpu = torch.empty(self.content_feat[0].shape)                   # shape carrier: indices are drawn on the host

        def feats(img):                      # img: (1, h, w, 3); its NCHW view is channels_last already
            return [img] + [nhwc(f) for f in vgg_feats(img.permute(0, 3, 1, 2))]

        def iteration(update):
            img = S.fold_laplacian_pyramid(self.variables)
            pred = feats(img)
            c_feat = sampling._sample(self.content_feat, self.static_idx, True)
            p_feat = sampling._sample(pred, self.static_idx, True)
            loss = self.loss_fn(c_feat, p_feat)
            grads = torch.autograd.grad(loss, self.variables)
            if update:
                self.opt.apply_gradients(zip(grads, self.variables))
            return loss

        # one eager forward + backward without an update, so that cuDNN plans/workspaces and the library workspace exist
        # before capture (no device allocation is allowed inside a capture)
        iteration(False)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = iteration(True)

    def load(self, content_feat, style_samples, stylized_nhwc):
        """A new image at this scale: refill the graph's buffers (same shapes), reset the optimizer slots."""
        with torch.no_grad():
            for d, src in zip(self.content_feat, content_feat):
                d.copy_(src)
            for v, src in zip(self.variables, S.make_laplacian_pyramid(stylized_nhwc, 5)):
                v.copy_(src)
            for slot in self.opt.slots():
                slot.zero_()
        self.loss_fn.handle.set_style_target(style_samples)

    def run(self, max_iter):
        nxt, last = None, None
        for it in range(max_iter):
            if nxt is not None:
                self.static_idx.copy_(nxt, non_blocking=True)
            self.graph.replay()
            nxt = self.sampling._make_indices(self.base_cpu, True).pin_memory()   # next iteration's draw, under the GPU work
            last = float(self.static_loss.item())   # the reference formats three scalars per iteration (run_strotss.py:150-152)
        return last

    def image(self):
        with torch.no_grad():
            return S.fold_laplacian_pyramid([v.detach() for v in self.variables])


_SCALE_GRAPHS = {}


def _run(args, content, style, sampling, feats_nchw, vgg_feats, max_iter):
    alpha = 16.0
    per_scale = []
    stylized = None
    for i in range(args.level):
        scl = 2 << (5 + i)
        torch.cuda.synchronize()
        t_setup = time.perf_counter()
        sc, ss = resize_long(content, scl), resize_long(style, scl)
        lap = make_laplacian(sc)
        lr = args.lr
        if i == 0:
            stylized = lap + ss.mean(dim=(2, 3), keepdim=True)
        elif i < args.level - 1:
            stylized = F.interpolate(stylized, size=sc.shape[-2:], mode="bilinear", align_corners=False) + lap
        else:
            stylized = F.interpolate(stylized, size=sc.shape[-2:], mode="bilinear", align_corners=False)
            lr = args.lr / 2
        feats, fold_pyramid = feats_nchw, fold_pyramid_torch
        with torch.no_grad():
            content_feat = feats(sc)
            style_samples = sampling(feats(ss))
        captured = False
        if args.eager:
            variables = [torch.nn.Parameter(v.clone()) for v in make_pyramid(stylized)]
            loss_fn = S.StrotssLoss(style_samples, alpha)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_vgg = t_loss = 0.0
        last = None
        if args.eager:
            opt = torch.optim.RMSprop(variables, lr=lr, alpha=0.99, eps=1e-8)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for it in range(max_iter):
                opt.zero_grad(set_to_none=True)
                ev[0].record()
                img = fold_pyramid(variables)
                pred = feats(img)
                ev[1].record()
                c_feat, p_feat = sampling.bilinear(content_feat, pred)
                loss = loss_fn(c_feat, p_feat)
                ev[2].record()
                loss.backward()
                opt.step()
                ev[3].record()
                last = float(loss.item())       # the reference formats three scalars per iteration (run_strotss.py:150-152)
                t_vgg += ev[0].elapsed_time(ev[1])
                t_loss += ev[1].elapsed_time(ev[2])
        else:
            key = (i, tuple(sc.shape), args.sample)
            t0 = time.perf_counter()
            runner = _SCALE_GRAPHS.get(key)
            if runner is None:
                runner = ScaleGraph(sampling, vgg_feats, content_feat, style_samples, nhwc(stylized), alpha, lr)
                _SCALE_GRAPHS[key] = runner
                captured = True
            else:
                runner.load(content_feat, style_samples, nhwc(stylized))
            torch.cuda.synchronize()
            t_warm, t_capture = 0.0, time.perf_counter() - t0
            last = runner.run(max_iter)
            variables = None
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        with torch.no_grad():
            if args.eager:
                stylized = fold_pyramid(variables).detach()
            else:
                stylized = runner.image().permute(0, 3, 1, 2)
        per_scale.append({"scale": scl, "content_hw": list(sc.shape[-2:]), "style_hw": list(ss.shape[-2:]), "alpha": alpha,
                          "seconds": dt, "ms_per_iter": dt / max_iter * 1e3, "setup_seconds": t0 - t_setup,
                          "graph_setup_seconds": (t_capture if not args.eager else 0.0),
                          "graph_captured_in_this_pass": (captured if not args.eager else False),
                          "fold_vgg_fwd_ms_per_iter": t_vgg / max_iter,
                          "sample_loss_fwd_ms_per_iter": t_loss / max_iter, "last_loss": last})
        alpha /= 2.0
    return per_scale


if __name__ == "__main__":
    main()
