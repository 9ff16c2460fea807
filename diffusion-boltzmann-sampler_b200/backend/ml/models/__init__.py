from .score_network import ScoreNetwork, SinusoidalTimeEmbedding, ConvBlock, SelfAttention
from .score_network_continuous import ContinuousScoreNetwork, PeriodicPositionEmbedding, ParticleAttention
from .diffusion import DiffusionProcess

__all__ = ["ScoreNetwork", "SinusoidalTimeEmbedding", "ConvBlock", "SelfAttention", "ContinuousScoreNetwork",
           "PeriodicPositionEmbedding", "ParticleAttention", "DiffusionProcess"]
