from .score_network import ScoreNetwork, SinusoidalTimeEmbedding, ConvBlock, SelfAttention
from .diffusion import DiffusionProcess

__all__ = ["ScoreNetwork", "SinusoidalTimeEmbedding", "ConvBlock", "SelfAttention", "DiffusionProcess"]
