"""CPU experiment (no GPU): what precision layouts can reach -- oracle U-Net under emulated roundings, rel-L2 vs the fp32 run.
  A: bf16 ONLY at matmul/conv INPUTS (weights + activations rounded, fp32 accumulate, fp32 everywhere else incl. residual stream)
  B: A + every module OUTPUT rounded to bf16 (the current CUDA design: bf16 between kernels)"""
import sys, torch, contextlib
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.nn.functional as F
from oracle import harness as helpers
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_full import full_inputs

MODE = {"round": "bf16"}


def bf(x):
    """round a tensor to the operand precision under test: bf16 (8 mantissa bits) or tf32 (11)"""
    if not (torch.is_tensor(x) and x.is_floating_point()):
        return x
    if MODE["round"] == "tf32":
        i = x.float().contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    return x.to(torch.bfloat16).float()

@contextlib.contextmanager
def matmul_inputs_bf16(round_outputs=False):
    saved = {}
    def wrap(mod, name, nin):
        orig = getattr(mod, name); saved[(mod, name)] = orig
        def f(*a, **k):
            a = list(a)
            for i in range(min(nin, len(a))): a[i] = bf(a[i])
            out = orig(*a, **k)
            return bf(out) if round_outputs else out
        setattr(mod, name, f)
    wrap(F, "linear", 2); wrap(F, "conv2d", 2); wrap(torch, "bmm", 2); wrap(torch, "matmul", 2)
    orig_baddbmm = torch.baddbmm; saved[(torch, "baddbmm")] = orig_baddbmm
    def baddbmm(inp, b1, b2, **k):
        out = orig_baddbmm(inp, bf(b1), bf(b2), **k); return bf(out) if round_outputs else out
    torch.baddbmm = baddbmm
    if round_outputs:
        for name in ("layer_norm", "group_norm", "silu", "gelu", "softmax"):
            orig = getattr(F, name); saved[(F, name)] = orig
            setattr(F, name, (lambda o: (lambda *a, **k: bf(o(*a, **k))))(orig))
    try: yield
    finally:
        for (mod, name), orig in saved.items(): setattr(mod, name, orig)

def rel(a, b): return float((a - b).norm() / b.norm())
for tiny in (True, False):
    u = helpers.build_oracle_unet(tiny=tiny, obj=True)
    if tiny:
        g = torch.Generator().manual_seed(1)
        r = lambda *s: bf(torch.randn(*s, generator=g))
        sample, text = r(2, 4, 8, 16, 24), 0.5 * r(2, 77, 768)
        feats = [r(2, C, 8, 16 >> l, 24 >> l) for l, C in enumerate((320, 640))]
        trajs = [0.5 * r(2, C, 8, 16 >> l, 24 >> l) for l, C in enumerate((320, 640))]
    else:
        inp = full_inputs(); sample, text, feats, trajs = bf(inp["sample"]), bf(inp["text"]), [bf(x) for x in inp["pose_feats"]], [bf(x) for x in inp["traj_feats"]]
    with torch.no_grad():
        want = u(sample, 961, text, pose_embedding_features=feats, traj_features=trajs).sample
        with matmul_inputs_bf16(False):
            a = u(sample, 961, text, pose_embedding_features=feats, traj_features=trajs).sample
        with matmul_inputs_bf16(True):
            b = u(sample, 961, text, pose_embedding_features=feats, traj_features=trajs).sample
        with torch.autocast("cpu", dtype=torch.bfloat16):
            c = u(sample, 961, text, pose_embedding_features=feats, traj_features=trajs).sample.float()
        MODE["round"] = "tf32"
        with matmul_inputs_bf16(False):
            d = u(sample, 961, text, pose_embedding_features=feats, traj_features=trajs).sample
        MODE["round"] = "bf16"
    print(f"{'tiny' if tiny else 'full'} U-Net: A (bf16 at matmul inputs only, fp32 residual stream) {rel(a, want):.2e} | "
          f"B (+ bf16 outputs of every op) {rel(b, want):.2e} | torch autocast bf16 {rel(c, want):.2e} | "
          f"tf32 at matmul inputs only {rel(d, want):.2e}", flush=True)
