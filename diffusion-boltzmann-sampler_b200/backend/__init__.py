"""Drop-in replacement modules for the hot path of Sakeeb91/diffusion-boltzmann-sampler.

Only `backend.ml.{systems,samplers,models,analysis}` exist here; the FastAPI app, training code
and front end of the reference are untouched consumers (INTEGRATION.md shows the overlay).
"""
__version__ = "1.0.0+b200"
