import sys, torch
sys.path.insert(0, '.')
import bench
from oracle import torch_port as T
import strotss_tensorflow_b200 as S
from strotss_tensorflow_b200 import _lib
dev = torch.device('cuda', 0)
for eps in (1.0, 0.1, 0.01):
    style, content, pred = bench.synth_torch(16384, 16384, 2179, eps, 0, dev)
    h = S.Handle(dev); h.set_style_target(style)
    sc, grad, _, _ = h.eval(pred, content, 16.0, True)
    s = sc.double().cpu().numpy(); g = grad.double(); del h
    ref, gref, info = T.total_loss_and_grad(style.double(), content.double(), pred.double(), 16.0)
    rel = {k: abs(s[slot] - info[k].item()) / info[k].item() for slot, k in [(_lib.S_LOSS_C, 'loss_c'), (_lib.S_LOSS_S, 'loss_s'), (_lib.S_L_M, 'l_m'), (_lib.S_L_REMD, 'l_remd'), (_lib.S_L_PALETTE, 'l_palette')]}
    gn, rn = g.norm().item(), gref.norm().item()
    print(f"eps={eps}: total rel {abs(s[0]-ref.item())/ref.item():.2e} " + " ".join(f"{k} {v:.1e}" for k, v in rel.items()) +
          f" | grad-norm rel {abs(gn-rn)/rn:.2e} cos {(g*gref).sum().item()/(gn*rn):.6f}", flush=True)
    del g, gref, ref, info; torch.cuda.empty_cache()
