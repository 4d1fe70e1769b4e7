"""Pins the two pipeline-edge restatements (oracle/vae.py, oracle/clip_text.py) to independent
implementations that ARE installed: transformers' own CLIPTextModel (the very dependency the reference calls) and
torchtitan's copy of the CompVis/LDM autoencoder (the architecture diffusers' AutoencoderKL ports)."""
import pytest
import torch


def test_clip_text_restatement_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from oracle.clip_text import CLIPTextModel as Oracle
    cfg = transformers.CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3,
                                      num_attention_heads=4, max_position_embeddings=77, hidden_act="quick_gelu")
    torch.manual_seed(0)
    ref = transformers.CLIPTextModel(cfg).eval()
    with torch.no_grad():  # the default init leaves LayerNorm at identity and biases at zero: move everything
        for p in ref.parameters():
            p.add_(0.05 * torch.randn_like(p))
    ora = Oracle(1000, 77, 128, 3, 4, 512, cfg.layer_norm_eps)
    state = {k: v for k, v in ref.state_dict().items() if "position_ids" not in k}
    missing, unexpected = ora.load_state_dict(state, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    ids = torch.randint(0, 1000, (2, 77), generator=torch.Generator().manual_seed(1))
    ids[:, -1] = 999  # eos = highest id (the pooled output is not used, but transformers locates it)
    with torch.no_grad():
        want = ref(ids)[0]
        got = ora(ids)[0]
    assert float((got - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))


def test_vae_restatement_matches_the_ldm_autoencoder():
    ae = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from oracle.vae import AutoencoderKL, ldm_state_to_diffusers
    ch, mult, z = 32, [1, 2, 4, 4], 4
    torch.manual_seed(0)
    enc = ae.Encoder(resolution=32, in_channels=3, ch=ch, ch_mult=mult, num_res_blocks=2, z_channels=z).eval()
    dec = ae.Decoder(ch=ch, out_ch=3, ch_mult=mult, num_res_blocks=2, in_channels=3, resolution=32, z_channels=z).eval()
    with torch.no_grad():
        for m in (enc, dec):
            for p in m.parameters():
                p.add_(0.05 * torch.randn_like(p))
    vae = AutoencoderKL(block_out_channels=tuple(ch * m for m in mult), latent_channels=z).eval()
    state = {f"encoder.{k}": v for k, v in enc.state_dict().items()}
    state.update({f"decoder.{k}": v for k, v in dec.state_dict().items()})
    missing, unexpected = vae.load_state_dict(ldm_state_to_diffusers(state), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("quant_conv", "post_quant_conv")) for k in missing), missing
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 40, 24, generator=g)
    zin = torch.randn(2, z, 5, 3, generator=g)
    with torch.no_grad():
        assert float((vae.encoder(x) - enc(x)).abs().max()) < 1e-5 * max(1.0, float(enc(x).abs().max()))
        assert float((vae.decoder(zin) - dec(zin)).abs().max()) < 1e-5 * max(1.0, float(dec(zin).abs().max()))


def test_vae_gaussian_posterior():
    from oracle.vae import DiagonalGaussianDistribution
    m = torch.randn(1, 8, 3, 3, generator=torch.Generator().manual_seed(3)) * 40
    d = DiagonalGaussianDistribution(m)
    assert float(d.logvar.max()) <= 20.0 and float(d.logvar.min()) >= -30.0
    noise = torch.randn(1, 4, 3, 3, generator=torch.Generator().manual_seed(4))
    assert torch.equal(d.sample(noise=noise), d.mean + torch.exp(0.5 * d.logvar) * noise)


def test_unet_diffusers_layer_restatement_matches_the_ldm_blocks():
    """The diffusers layer UNDER the U-Net oracle (oracle/diffusers_restated.py: ResnetBlock2D, Attention, Upsample2D) has no
    reference fixtures -- diffusers is not installed.  Where those blocks coincide with the CompVis/LDM blocks diffusers
    ported them from (installed through torchtitan), they are held to them here: the residual block with the
    time-embedding term switched off, single-head attention with biases (both restated processors: SDPA and the explicit
    baddbmm / softmax / bmm path of fmc's AttnProcessor), and the nearest-2x upsampler."""
    ae = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from oracle import diffusers_restated as dr
    from oracle.attention_processor import AttnProcessor
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)

    def jitter(m):
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.05 * torch.randn_like(p))
        return m.eval()
    for cin, cout in ((64, 64), (64, 128)):
        ref = jitter(ae.ResnetBlock(cin, cout))
        blk = dr.ResnetBlock2D(cin, cout, temb_channels=32, groups=32, eps=1e-6).eval()
        state = {k.replace("nin_shortcut", "conv_shortcut"): v for k, v in ref.state_dict().items()}
        state["time_emb_proj.weight"] = torch.zeros(cout, 32)   # no time embedding in the LDM block
        state["time_emb_proj.bias"] = torch.zeros(cout)
        blk.load_state_dict(state, strict=True)
        x = torch.randn(2, cin, 12, 10, generator=g)
        with torch.no_grad():
            want, got = ref(x), blk(x, torch.randn(2, 32, generator=g))
        assert float((got - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))
    C = 64
    ref = jitter(ae.AttnBlock(C))
    x = torch.randn(2, C, 6, 5, generator=g)
    for processor in (dr.AttnProcessorSDPA(), AttnProcessor()):
        attn = dr.Attention(C, heads=1, dim_head=C, bias=True, processor=processor).eval()
        sd = ref.state_dict()
        attn.load_state_dict({"to_q.weight": sd["q.weight"][:, :, 0, 0], "to_q.bias": sd["q.bias"],
                              "to_k.weight": sd["k.weight"][:, :, 0, 0], "to_k.bias": sd["k.bias"],
                              "to_v.weight": sd["v.weight"][:, :, 0, 0], "to_v.bias": sd["v.bias"],
                              "to_out.0.weight": sd["proj_out.weight"][:, :, 0, 0], "to_out.0.bias": sd["proj_out.bias"]},
                             strict=True)
        with torch.no_grad():
            tokens = ref.norm(x).flatten(2).transpose(1, 2)                       # [b, h w, C]
            got = x + attn(tokens).transpose(1, 2).reshape(x.shape)
            want = ref(x)
        assert float((got - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max())), type(processor).__name__
    ref = jitter(ae.Upsample(C))
    up = dr.Upsample2D(C).eval()
    up.load_state_dict(ref.state_dict(), strict=True)
    with torch.no_grad():
        assert float((up(x) - ref(x)).abs().max()) < 1e-6


def test_timestep_embedding_restatement_matches_an_independent_implementation():
    """diffusers `Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0)` + `TimestepEmbedding` (unet.py:1090-1096) as
    restated in oracle/diffusers_restated.py against the sinusoidal embedding + two-layer SiLU MLP of torchtitan's flux
    model (the same published formula, written independently: [cos | sin] of t * 10000^(-i / half))."""
    layers = pytest.importorskip("torchtitan.experiments.flux.model.layers")
    from oracle import diffusers_restated as dr
    t = torch.tensor([961.0, 41.0, 1.0, 500.0])
    want = layers.timestep_embedding(t, 320, time_factor=1.0)
    got = dr.Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0)(t)
    assert got.shape == (4, 320) and float((got - want).abs().max()) < 1e-6
    torch.manual_seed(0)
    ref = layers.MLPEmbedder(320, 1280).eval()
    emb = dr.TimestepEmbedding(320, 1280).eval()
    emb.load_state_dict({"linear_1.weight": ref.in_layer.weight, "linear_1.bias": ref.in_layer.bias,
                         "linear_2.weight": ref.out_layer.weight, "linear_2.bias": ref.out_layer.bias}, strict=True)
    with torch.no_grad():
        assert float((emb(got) - ref(want)).abs().max()) < 1e-5
