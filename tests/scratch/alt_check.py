import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import strotss_oracle as O
import strotss_tensorflow_b200 as S
dev = torch.device('cuda', 0)
for (N, M, eps) in [(2300, 700, 0.1), (4500, 600, 0.01)]:
    st, co, pr = O.synth_problem(N, M, 2179, eps=eps, seed=41)
    mod = S.StrotssLoss(torch.tensor(st, device=dev), 16.0)
    sc, grad, _, _ = mod.handle.eval(torch.tensor(pr, device=dev), torch.tensor(co, device=dev), 16.0, True)
    ref, gref, info = O.total_loss(st, co, pr, 16.0, np.float64, True)
    g = grad.double().cpu().numpy()
    print('REL', N, abs(sc[0].item() - ref) / ref, abs(np.linalg.norm(g) - np.linalg.norm(gref)) / np.linalg.norm(gref),
      float((g * gref).sum() / (np.linalg.norm(g) * np.linalg.norm(gref))), flush=True)
