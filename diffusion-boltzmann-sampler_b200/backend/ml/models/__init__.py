"""Score networks, the VP-SDE and the noise schedules; exports the names of the reference's models package."""
from . import diffusion as _sde, noise_schedule as _ns, score_network as _unet, score_network_continuous as _csn

_EXPORTS = {
    _unet: ("ScoreNetwork", "SinusoidalTimeEmbedding", "ConvBlock", "SelfAttention"),
    _csn: ("ContinuousScoreNetwork", "PeriodicPositionEmbedding", "ParticleAttention"),
    _sde: ("DiffusionProcess",),
    _ns: ("NoiseSchedule", "LinearNoiseSchedule", "CosineNoiseSchedule", "SigmoidNoiseSchedule", "get_schedule"),
}
for _mod, _names in _EXPORTS.items():
    for _n in _names:
        globals()[_n] = getattr(_mod, _n)

__all__ = tuple(n for names in _EXPORTS.values() for n in names)
