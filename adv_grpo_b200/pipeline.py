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
        self.transformer = transformer
        self.vae = vae
        self.scheduler = scheduler or FlowMatchEulerDiscreteScheduler()
        self.image_processor = VaeImageProcessor()
        self.tokenizer = tokenizer or SyntheticCLIPTokenizer()
        self._execution_device = torch.device(device)
        self.default_sample_size = transformer.config.sample_size
        self.graphed_transformer = GraphedForward(transformer) if use_cuda_graph else None

    @classmethod
    def from_seed(cls, mmdit_cfg=weights.SD35_MEDIUM, vae_cfg=weights.VAE_SD3, device="cuda", lora_rank=32,
                  lora_alpha=64, seed=0, lora=None, use_cuda_graph=True):
        """Seeded random weights at the true shapes (no checkpoints exist on the box)."""
        tp = weights.init_mmdit(mmdit_cfg, seed=seed, device=device, dtype=torch.bfloat16)
        transformer = SD3Transformer2DModel(mmdit_cfg, tp, lora_rank=lora_rank, lora_alpha=lora_alpha, lora=lora,
                                            device=device)
        vae = AutoencoderKL(weights.init_vae_decoder(vae_cfg, seed=seed + 2, device=device), vae_cfg, device=device)
        return cls(transformer, vae, device=device, use_cuda_graph=use_cuda_graph)
