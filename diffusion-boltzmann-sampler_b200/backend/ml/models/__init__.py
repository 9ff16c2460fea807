from .score_network import ScoreNetwork, SinusoidalTimeEmbedding, ConvBlock, SelfAttention
from .score_network_continuous import ContinuousScoreNetwork, PeriodicPositionEmbedding, ParticleAttention
from .diffusion import DiffusionProcess
from .noise_schedule import (NoiseSchedule, LinearNoiseSchedule, CosineNoiseSchedule, SigmoidNoiseSchedule,
                             get_schedule)

__all__ = ["ScoreNetwork", "SinusoidalTimeEmbedding", "ConvBlock", "SelfAttention", "ContinuousScoreNetwork",
           "PeriodicPositionEmbedding", "ParticleAttention", "DiffusionProcess", "NoiseSchedule",
           "LinearNoiseSchedule", "CosineNoiseSchedule", "SigmoidNoiseSchedule", "get_schedule"]
