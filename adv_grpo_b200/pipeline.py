"""`StableDiffusion3Pipeline`-shaped container for the objects the hot path reaches through `pipeline.*`
(`transformer`, `scheduler`, `vae`, `image_processor`, `tokenizer`) plus a CUDA-graph
runner for the rollout forward.  Text encoders are outside the path (prompt embeddings are inputs)."""
import torch

from . import _lib, weights
from .mmdit import SD3Transformer2DModel
from .pickscore_scorer import SyntheticCLIPTokenizer
from .scheduler import FlowMatchEulerDiscreteScheduler
from .vae import AutoencoderKL, VaeImageProcessor


class GraphedForward:
    """Replays one captured MMDiT forward per input-shape key (no_grad rollout only).  The captured kernels read the
    model's persistent bf16 LoRA operand buffers; those are refreshed IN PLACE (eagerly, before the replay) whenever
    `invalidate_lora_cache()` marked them dirty (optimizer step, EMA swap, `load_adapter`), so a graph stays valid
    across parameter updates and never replays stale weights."""

    def __init__(self, model, warmup=2):
        self.model, self.warmup, self.graphs = model, warmup, {}

    def __call__(self, hidden_states, timestep, encoder_hidden_states, pooled_projections, return_dict=False, **_):
        key = (tuple(hidden_states.shape), tuple(encoder_hidden_states.shape), self.model._lora_enabled)
        ent = self.graphs.get(key)
        if ent is None:
            static_in = [hidden_states.clone(), timestep.clone().float(), encoder_hidden_states.clone(),
                         pooled_projections.clone()]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(self.warmup):
                    self.model(*static_in)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.no_grad(), torch.cuda.graph(g):
                out = self.model(*static_in)[0]
            ent = (g, static_in, out, _lib.launch_count() - n0)
            self.graphs[key] = ent
        g, static_in, out, n_kernels = ent
        with torch.no_grad():
            self.model._pack_lora()          # no-op unless dirty; the repack writes into the buffers the graph reads
        static_in[0].copy_(hidden_states)
        static_in[1].copy_(timestep)
        static_in[2].copy_(encoder_hidden_states)
        static_in[3].copy_(pooled_projections)
        g.replay()
        _lib.add_launches(n_kernels)
        return (out,)


class StableDiffusion3Pipeline:
    vae_scale_factor = 8

    def __init__(self, transformer, vae, scheduler=None, tokenizer=None, device="cuda", use_cuda_graph=True):
        self._use_cuda_graph = use_cuda_graph
        self.transformer = transformer
        self.vae = vae
        self.scheduler = scheduler or FlowMatchEulerDiscreteScheduler()
        self.image_processor = VaeImageProcessor()
        self.tokenizer = tokenizer or SyntheticCLIPTokenizer()
        self._execution_device = torch.device(device)
        self.default_sample_size = transformer.config.sample_size
        self._use_cuda_graph = use_cuda_graph
        self.text_encoder = self.text_encoder_2 = self.text_encoder_3 = None
        self.tokenizer_2 = self.tokenizer_3 = None
        self.safety_checker = None

    def set_progress_bar_config(self, **kw):
        pass

    @property
    def transformer(self):
        return self._transformer

    @transformer.setter
    def transformer(self, model):
        """`pipeline.transformer = get_peft_model(pipeline.transformer, ...)` (train_pick:506-511): the rollout graph
        runner follows the new object."""
        self._transformer = model
        self.graphed_transformer = GraphedForward(model) if getattr(self, "_use_cuda_graph", False) else None

    @classmethod
    def from_seed(cls, mmdit_cfg=weights.SD35_MEDIUM, vae_cfg=weights.VAE_SD3, device="cuda", lora_rank=32,
                  lora_alpha=64, seed=0, lora=None, use_cuda_graph=True):
        """Seeded random weights at the true shapes (no checkpoints exist on the box)."""
        tp = weights.init_mmdit(mmdit_cfg, seed=seed, device=device, dtype=torch.bfloat16)
        transformer = SD3Transformer2DModel(mmdit_cfg, tp, lora_rank=lora_rank, lora_alpha=lora_alpha, lora=lora,
                                            device=device)
        vae = AutoencoderKL(weights.init_vae_decoder(vae_cfg, seed=seed + 2, device=device), vae_cfg, device=device)
        return cls(transformer, vae, device=device, use_cuda_graph=use_cuda_graph)

    def add_seeded_text_encoders(self, clip_l=weights.CLIP_L_TEXT, clip_g=weights.CLIP_G_TEXT, t5=weights.T5_XXL, seed=5):
        """CLIP-L / CLIP-G / T5 encoders with seeded weights at the given sizes + synthetic tokenizers (no checkpoints on
        the box), so `pipeline.text_encoder{,_2,_3}` / `pipeline.tokenizer{,_2,_3}` of train_pick:456-457 exist."""
        from . import text_encoders as te
        dev = self._execution_device
        self.text_encoder = te.CLIPTextModelWithProjection(weights.init_clip_text(clip_l, seed=seed, device=dev), clip_l, device=dev)
        self.text_encoder_2 = te.CLIPTextModelWithProjection(weights.init_clip_text(clip_g, seed=seed + 1, device=dev), clip_g, device=dev)
        self.text_encoder_3 = te.T5EncoderModel(weights.init_t5_encoder(t5, seed=seed + 2, device=dev), t5, device=dev)
        self.tokenizer = SyntheticCLIPTokenizer(clip_l["vocab"])
        self.tokenizer_2 = SyntheticCLIPTokenizer(clip_g["vocab"])
        self.tokenizer_3 = SyntheticCLIPTokenizer(t5["vocab"])
        return self
