import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200 import losses, ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
dev = torch.device("cuda:0")
B, H, W, S = 1, 256, 832, 3
t = make_triplet(B, H, W, 4, S, seed=71, flow_mode="rigid")
loss_o, aux = P.depth_mode_loss(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K, S, "live", return_aux=True)
td = t.to(dev)
loss, masks = losses.DepthLoss(S, "live").forward_losses(td.img_l, td.img, td.img_r, td.disp, td.disp_l, td.disp_r, td.pose, td.K)
pc = P.bilinear_pyramid(t.img, S); pl = P.bilinear_pyramid(t.img_l, S); pr = P.bilinear_pyramid(t.img_r, S)
for mk in ("valid_l", "valid_r", "tex_b", "tex_f"):
    for l in range(S):
        d = (masks[mk][l].cpu() != aux[mk][l])
        if d.any():
            idx = d.nonzero()
            print(mk, l, idx.tolist())
            for b, c, i, j in idx.tolist():
                rec = aux["rec_l" if mk.endswith("b") or mk.endswith("_l") else "rec_r"][l]
                src = (pl if mk.endswith("b") else pr)[l]
                a = (pc[l][b, :, i, j] - rec[b, :, i, j]).abs().mean().item()
                s = (pc[l][b, :, i, j] - src[b, :, i, j]).abs().mean().item()
                print("  margins: mean|I-rec| %.9g  mean|I-src| %.9g  diff %.3g" % (a, s, a - s))
# matrices: kernel vs torch CPU
Kinv, (P_b, P_f), _ = ops.pose_setup(td.pose, td.K, [1.0, 2.0, 4.0])
for s, ds in enumerate([1.0, 2.0, 4.0]):
    Ks = P.scale_intrinsics(t.K, ds)
    print("level", s, "Kinv maxdiff", (Kinv[s].cpu() - torch.inverse(Ks)).abs().max().item(), "P_b maxdiff", (P_b[s].cpu() - Ks.bmm(P.pose_to_matrix(t.pose[:, 0]))).abs().max().item(),
          "P_f", (P_f[s].cpu() - Ks.bmm(P.pose_to_matrix(t.pose[:, 1]))).abs().max().item())
