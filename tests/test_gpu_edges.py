"""the B200 VAE (synfmc_b200/edge/autoencoder_kl.py) and CLIP text encoder (edge/clip_text.py) against their
fp32 CPU restatements (oracle/vae.py, oracle/clip_text.py -- themselves pinned to torchtitan's LDM autoencoder and to
transformers in tests/test_oracle_edges.py), on the same seeded weights (bf16-representable) and inputs.

Tolerances (bf16 activations, fp32 accumulation; written here, measured values printed): kernels 4e-3 rel-L2; the text
encoder (12 layers) 1.5e-2; VAE decode / encode (about 30 / 25 convolution layers) 2e-2."""
import pytest
import torch

from oracle.harness import rel_l2

pytestmark = pytest.mark.gpu

KERNEL_TOL, CLIP_TOL, VAE_TOL = 4e-3, 1.5e-2, 2e-2


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _pair_vae(cuda_device, ch, seed=11):
    from oracle.vae import AutoencoderKL as OV
    from synfmc_b200.edge import AutoencoderKL
    from synfmc_b200.synth import synth_init_
    ov = OV(block_out_channels=ch).eval().requires_grad_(False)
    synth_init_(ov, seed=seed)
    pv = AutoencoderKL(block_out_channels=ch)
    assert pv.load_state_dict(ov.state_dict(), strict=True)
    return ov, pv.to(cuda_device)


def test_softmax_rows(cuda_device):
    from synfmc_b200 import ops
    g = torch.Generator().manual_seed(0)
    for rows, n in ((7, 64), (33, 2560), (5, 4096)):
        s = torch.randn(rows, n, generator=g) * 20
        want = torch.softmax(s * 0.0442, dim=-1)
        with torch.cuda.device(cuda_device):
            got32 = ops.softmax_rows(s.to(cuda_device), 0.0442, out_dtype=torch.float32).cpu()
            got16 = ops.softmax_rows(s.to(cuda_device), 0.0442).float().cpu()
        assert rel_l2(got32, want) < 1e-5 and rel_l2(got16, want) < KERNEL_TOL


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("d,T,heads", [(64, 77, 12), (32, 20, 4), (128, 128, 2)])
def test_small_mha(cuda_device, causal, d, T, heads):
    from synfmc_b200 import ops
    B, C = 3, heads * d
    qkv = _bf(torch.randn(B * T, 3 * C, generator=torch.Generator().manual_seed(d + T)))
    q, k, v = (qkv[:, i * C:(i + 1) * C].view(B, T, heads, d).transpose(1, 2).double() for i in range(3))
    want = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(B * T, C)
    with torch.cuda.device(cuda_device):
        got16 = ops.small_mha(qkv.to(cuda_device).bfloat16(), 0, C, 2 * C, B, T, heads, d, d ** -0.5, causal).float().cpu()
        got32 = ops.small_mha(qkv.to(cuda_device), 0, C, 2 * C, B, T, heads, d, d ** -0.5, causal).cpu()
    assert rel_l2(got32, want) < 1e-5 and rel_l2(got16, want) < KERNEL_TOL


def test_quick_gelu_embed_sample_video_kernels(cuda_device):
    from synfmc_b200 import ops
    g = torch.Generator().manual_seed(3)
    dev = cuda_device
    with torch.cuda.device(dev):
        x = _bf(torch.randn(50, 96, generator=g) * 3)
        assert rel_l2(ops.quick_gelu(x.to(dev)).cpu(), x * torch.sigmoid(1.702 * x)) < 1e-6
        assert rel_l2(ops.quick_gelu(x.to(dev).bfloat16()).float().cpu(), x * torch.sigmoid(1.702 * x)) < KERNEL_TOL
        tok, pos = torch.randn(100, 64, generator=g), torch.randn(77, 64, generator=g)
        ids = torch.randint(0, 100, (2, 77), generator=g)
        got = ops.embed_tokens(ids.to(dev), tok.to(dev), pos.to(dev), out_dtype=torch.float32).cpu()
        assert torch.equal(got, (tok[ids] + pos[None, :77]).view(-1, 64))
        N, z, HW = 2, 4, 48
        m = torch.randn(N * HW, 2 * z, generator=g) * 25          # logvar beyond both clamps
        noise = torch.randn(N, z, HW, generator=g)
        mean = m[:, :z].view(N, HW, z).permute(0, 2, 1)
        logvar = m[:, z:].view(N, HW, z).permute(0, 2, 1).clamp(-30, 20)
        want = (mean + torch.exp(0.5 * logvar) * noise) * 0.18215
        got = ops.vae_sample(m.to(dev), N, z, HW, noise.to(dev), 0.18215).cpu()
        assert rel_l2(got, want) < 1e-6
        assert torch.equal(ops.vae_sample(m.to(dev), N, z, HW).cpu(), mean.contiguous())
        y = torch.randn(2 * 3 * 20, 8, generator=g) * 2            # [(b f) HW, ld >= C]
        want = (y[:, :3].view(2, 3, 20, 3).permute(0, 3, 1, 2) / 2 + 0.5).clamp(0, 1)
        assert torch.allclose(ops.cl_to_video(y.to(dev), 2, 3, 3, 20, 0.5, 0.5, 0.0, 1.0).cpu(), want, atol=1e-7)


def test_clip_text_encoder_sd15_shape(cuda_device):
    """The SD1.5 text tower (12 x 768, 12 heads, 77 tokens) on a CFG pair of prompts: last_hidden_state against the
    restatement."""
    from oracle.clip_text import CLIPTextModel as OC
    from synfmc_b200.edge import CLIPTextModel
    from synfmc_b200.synth import synth_init_
    oc = OC().eval().requires_grad_(False)
    synth_init_(oc, seed=5)
    pc = CLIPTextModel()
    assert pc.load_state_dict(oc.state_dict(), strict=True)
    pc.to(cuda_device)
    ids = torch.randint(0, 49408, (2, 77), generator=torch.Generator().manual_seed(6))
    want = oc(ids)[0]
    got = pc(ids.to(cuda_device))
    assert got[0].shape == (2, 77, 768) and got[0].dtype == torch.float32 and got.last_hidden_state is got[0]
    e = rel_l2(got[0].cpu(), want)
    print(f"[parity] CLIP text encoder 12x768, 2x77 tokens: rel-L2 {e:.3e}")
    assert e < CLIP_TOL
    with pytest.raises(NotImplementedError):
        pc(ids.to(cuda_device), attention_mask=torch.ones(2, 77))


@pytest.mark.parametrize("ch,hw", [((32, 64, 128, 128), (8, 16)), ((128, 256, 512, 512), (8, 8))])
def test_vae_decode_and_encode(cuda_device, ch, hw):
    """decode(z).sample and encode(x).latent_dist (mode and a sample with a given noise draw), small and SD1.5 widths."""
    ov, pv = _pair_vae(cuda_device, ch)
    g = torch.Generator().manual_seed(7)
    h, w = hw
    z = torch.randn(2, 4, h, w, generator=g)
    want = ov.decode(z).sample
    got = pv.decode(z.to(cuda_device)).sample
    assert got.shape == want.shape == (2, 3, 8 * h, 8 * w)
    e_dec = rel_l2(got.cpu(), want)
    x = torch.rand(2, 3, 8 * h, 8 * w, generator=g) * 2 - 1
    od, pd = ov.encode(x).latent_dist, pv.encode(x.to(cuda_device)).latent_dist
    noise = torch.randn(2, 4, h, w, generator=g)
    e_mode = rel_l2(pd.mode().cpu(), od.mode())
    e_samp = rel_l2(pd.sample(noise=noise.to(cuda_device), scale=0.18215).cpu(), od.sample(noise=noise) * 0.18215)
    print(f"[parity] VAE {ch} latent {h}x{w}: decode {e_dec:.3e}, encode mode {e_mode:.3e}, sample {e_samp:.3e}")
    assert e_dec < VAE_TOL and e_mode < VAE_TOL and e_samp < VAE_TOL


def test_vae_decode_video_matches_reference_decode_latents(cuda_device):
    """decode_video == the reference's decode_latents (per-frame decode of latents / 0.18215, rearrange, /2 + 0.5, clamp)."""
    ov, pv = _pair_vae(cuda_device, (32, 64, 128, 128))
    lat = torch.randn(2, 4, 3, 8, 8, generator=torch.Generator().manual_seed(8)) * 0.18215 * 3
    frames = (lat / 0.18215).permute(0, 2, 1, 3, 4).reshape(6, 4, 8, 8)
    want = torch.cat([ov.decode(frames[i:i + 1]).sample for i in range(6)]).view(2, 3, 3, 64, 64).permute(0, 2, 1, 3, 4)
    want = (want / 2 + 0.5).clamp(0, 1)
    got = pv.decode_video(lat.to(cuda_device))
    assert got.shape == (2, 3, 3, 64, 64) and float(got.min()) >= 0.0 and float(got.max()) <= 1.0
    e = rel_l2(got.cpu(), want)
    print(f"[parity] VAE decode_video 2 clips x 3 frames: rel-L2 {e:.3e}")
    assert e < VAE_TOL
    assert torch.equal(pv.decode_video(lat.to(cuda_device), chunk=2), got)   # chunked decode: same frames, same kernels


@pytest.mark.timeout(600)
def test_vae_decode_full_size_frame(cuda_device):
    """One 320x512 frame (latent 40x64, 2560 attention tokens) of the SD1.5 decoder."""
    ov, pv = _pair_vae(cuda_device, (128, 256, 512, 512), seed=12)
    import time
    z = torch.randn(1, 4, 40, 64, generator=torch.Generator().manual_seed(9))
    t0 = time.perf_counter()
    want = ov.decode(z).sample
    cpu_ms = (time.perf_counter() - t0) * 1e3
    got = pv.decode(z.to(cuda_device)).sample
    e = rel_l2(got.cpu(), want)
    print(f"[parity] VAE decode 320x512 frame: rel-L2 {e:.3e} (fp32 restatement on {torch.get_num_threads()} host threads: "
          f"{cpu_ms:.0f} ms for the frame, first call)")
    assert e < VAE_TOL


class _StubTokenizer:
    """Stands in for transformers' CLIPTokenizer (its vocabulary files are not available offline): deterministic ids per
    prompt string, padded to 77 with the end-of-text id -- the call signature the pipelines use."""
    model_max_length = 77

    def __call__(self, texts, padding=None, max_length=77, truncation=True, return_tensors="pt"):
        texts = [texts] if isinstance(texts, str) else list(texts)
        ids = torch.full((len(texts), max_length), 49407, dtype=torch.long)
        for i, t in enumerate(texts):
            g = torch.Generator().manual_seed(sum(map(ord, t)) + 1)
            n = min(len(t.split()) + 2, max_length)
            ids[i, :n] = torch.randint(0, 49000, (n,), generator=g)
            ids[i, 0] = 49406
        return type("Enc", (), {"input_ids": ids})()


@pytest.mark.timeout(900)
def test_sample_call_end_to_end_with_the_b200_text_encoder_and_vae(cuda_device):
    """One whole `CameraCtrlPipeline.__call__` (fmc/pipelines/pipeline_animation.py:570-702): prompt + negative prompt ->
    text encoder -> CameraEncoder -> 3 CFG denoising steps -> VAE decode -> `.videos` [b, 3, f, H, W] in [0, 1] on the CPU,
    every stage on the library's kernels, against the composition of the CPU restatements."""
    from oracle.clip_text import CLIPTextModel as OC
    from oracle.diffusers_restated import DDIMScheduler as ODDIM
    from oracle.pipeline import denoise as o_denoise
    from oracle.rays import to_plucker_embedding
    from oracle.vae import AutoencoderKL as OV
    from oracle import harness as helpers
    from synfmc_b200 import synth
    from synfmc_b200.edge import AutoencoderKL, CLIPTextModel
    from synfmc_b200.fmc._blocks import DDIMScheduler
    from synfmc_b200.fmc.pipelines.pipeline_animation import CameraCtrlPipeline
    from synfmc_b200.synth import synth_init_
    dev = cuda_device
    channels = (320, 640)
    o_unet = helpers.build_oracle_unet(tiny=True)
    p_unet = helpers.build_product_unet(o_unet, tiny=True, device=dev)
    o_enc = helpers.build_oracle_pose_encoder(channels)
    p_enc = helpers.build_product_pose_encoder(o_enc, channels, device=dev)
    oc = OC(num_hidden_layers=2).eval().requires_grad_(False)              # SD1.5 width (768 = the U-Net's text width)
    synth_init_(oc, seed=21)
    pc = CLIPTextModel(num_hidden_layers=2)
    pc.load_state_dict(oc.state_dict())
    ov = OV(block_out_channels=(32, 64, 128, 128)).eval().requires_grad_(False)
    synth_init_(ov, seed=22)
    pv = AutoencoderKL(block_out_channels=(32, 64, 128, 128))
    pv.load_state_dict(ov.state_dict())
    tok = _StubTokenizer()
    b, f, H, W = 1, 8, 64, 128
    K, c2w = synth.synth_camera(b, f, H, W, seed=4)
    plucker = to_plucker_embedding(c2w, K, (H, W)).permute(0, 2, 1, 3, 4).contiguous()
    latents, _ = synth.synth_step_inputs(b, f, H // 8, W // 8, cfg=True, seed=4)
    prompt, negative = "a red car drives through a forest", "blurry"
    # restatement chain
    text = torch.cat([oc(tok([negative]).input_ids)[0], oc(tok([prompt]).input_ids)[0]])
    lat = o_denoise(o_unet, ODDIM(), o_enc, latents, text, plucker, f, num_inference_steps=25, guidance_scale=7.5, max_steps=3)
    frames = (lat / 0.18215).permute(0, 2, 1, 3, 4).reshape(b * f, 4, H // 8, W // 8)
    want = torch.cat([ov.decode(frames[i:i + 1]).sample for i in range(b * f)]).view(b, f, 3, H, W).permute(0, 2, 1, 3, 4)
    want = (want / 2 + 0.5).clamp(0, 1)
    # product
    pipe = CameraCtrlPipeline(pv.to(dev), pc.to(dev), tok, p_unet, DDIMScheduler(), p_enc)
    out = pipe(prompt, plucker.to(dev), f, height=H, width=W, num_inference_steps=25, guidance_scale=7.5,
               negative_prompt=negative, latents=latents.to(dev), max_steps=3)
    assert out.videos.shape == (b, 3, f, H, W) and out.videos.device.type == "cpu" and out.videos.dtype == torch.float32
    assert float(out.videos.min()) >= 0.0 and float(out.videos.max()) <= 1.0
    e_lat, e_vid = rel_l2(out.latents.cpu(), lat), rel_l2(out.videos, want)
    print(f"[parity] sample call end to end (text encoder -> 3 CFG steps -> VAE): latents {e_lat:.3e}, video {e_vid:.3e}")
    assert e_lat < 3e-2 and e_vid < 3e-2
