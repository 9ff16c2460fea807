"""Tests for DiffusionSampler."""

import pytest
import torch
import numpy as np
from backend.ml.samplers.diffusion import DiffusionSampler, PretrainedDiffusionSampler
from backend.ml.models.score_network import ScoreNetwork
from backend.ml.models.diffusion import DiffusionProcess


class TestDiffusionSamplerInit:
    """Tests for DiffusionSampler initialization."""

    def test_init_with_score_network(self):
        """Sampler should initialize with score network."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        sampler = DiffusionSampler(
            score_network=model,
            num_steps=50,
        )
        assert sampler.model is model
        assert sampler.num_steps == 50

    def test_init_with_custom_diffusion(self):
        """Sampler should accept custom diffusion process."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        diffusion = DiffusionProcess(beta_min=0.05, beta_max=15.0)
        sampler = DiffusionSampler(
            score_network=model,
            diffusion=diffusion,
            num_steps=50,
        )
        assert sampler.diffusion is diffusion

    def test_init_default_diffusion(self):
        """Sampler should create default diffusion if not provided."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        sampler = DiffusionSampler(score_network=model)
        assert sampler.diffusion is not None
        assert isinstance(sampler.diffusion, DiffusionProcess)


class TestDiffusionSamplerSample:
    """Tests for the sample method."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=20,
        )

    def test_sample_returns_tensor(self, diffusion_sampler):
        """Sample should return a tensor."""
        samples = diffusion_sampler.sample(shape=(2, 1, 8, 8))
        assert isinstance(samples, torch.Tensor)

    def test_sample_correct_shape(self, diffusion_sampler):
        """Sample should return correct shape."""
        samples = diffusion_sampler.sample(shape=(4, 1, 8, 8))
        assert samples.shape == (4, 1, 8, 8)

    def test_sample_batch_size_one(self, diffusion_sampler):
        """Sample should work with batch_size=1."""
        samples = diffusion_sampler.sample(shape=(1, 1, 8, 8))
        assert samples.shape == (1, 1, 8, 8)


class TestDiffusionSamplerDiscretize:
    """Tests for spin discretization methods."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_discretize_sign_method(self, diffusion_sampler):
        """Sign method should produce ±1 values."""
        continuous = torch.randn(2, 1, 8, 8)
        discrete = diffusion_sampler.discretize_spins(continuous, method="sign")
        unique = torch.unique(discrete)
        assert all(v in [-1, 1] for v in unique.tolist())

    def test_discretize_tanh_method(self, diffusion_sampler):
        """Tanh method should produce values in [-1, 1]."""
        continuous = torch.randn(2, 1, 8, 8) * 2
        discrete = diffusion_sampler.discretize_spins(continuous, method="tanh")
        assert (discrete >= -1).all()
        assert (discrete <= 1).all()

    def test_discretize_stochastic_method(self, diffusion_sampler):
        """Stochastic method should produce ±1 values."""
        continuous = torch.randn(2, 1, 8, 8)
        discrete = diffusion_sampler.discretize_spins(continuous, method="stochastic")
        unique = torch.unique(discrete)
        assert all(v in [-1, 1] for v in unique.tolist())

    def test_discretize_gumbel_method(self, diffusion_sampler):
        """Gumbel method should produce soft values in [-1, 1] range."""
        # Gumbel softmax produces soft values (expected value of P(+1) - P(-1))
        continuous = torch.randn(2, 1, 8, 8)
        discrete = diffusion_sampler.discretize_spins(continuous, method="gumbel")
        # Values should be in [-1, 1] range (soft discretization)
        assert (discrete >= -1).all()
        assert (discrete <= 1).all()

    def test_discretize_preserves_shape(self, diffusion_sampler):
        """Discretization should preserve tensor shape."""
        continuous = torch.randn(3, 1, 8, 8)
        discrete = diffusion_sampler.discretize_spins(continuous, method="sign")
        assert discrete.shape == continuous.shape


class TestDiffusionSamplerSampleIsing:
    """Tests for sample_ising method."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_sample_ising_returns_discrete(self, diffusion_sampler):
        """sample_ising should return discrete ±1 values."""
        samples = diffusion_sampler.sample_ising(batch_size=2, lattice_size=8)
        unique = torch.unique(samples)
        assert all(v in [-1, 1] for v in unique.tolist())

    def test_sample_ising_correct_shape(self, diffusion_sampler):
        """sample_ising should return correct shape without channel dimension."""
        samples = diffusion_sampler.sample_ising(batch_size=3, lattice_size=8)
        # sample_ising removes channel dimension, returns (batch, height, width)
        assert samples.shape == (3, 8, 8)


class TestDiffusionSamplerTrajectory:
    """Tests for sample_with_trajectory method."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=20,
        )

    def test_trajectory_is_generator(self, diffusion_sampler):
        """sample_with_trajectory should return a generator."""
        from types import GeneratorType
        trajectory = diffusion_sampler.sample_with_trajectory(shape=(1, 1, 8, 8))
        assert isinstance(trajectory, GeneratorType)

    def test_trajectory_yields_tuples(self, diffusion_sampler):
        """Trajectory should yield (tensor, time) tuples."""
        for x, t in diffusion_sampler.sample_with_trajectory(shape=(1, 1, 8, 8), yield_every=5):
            assert isinstance(x, torch.Tensor)
            assert isinstance(t, float)
            break

    def test_trajectory_time_decreases(self, diffusion_sampler):
        """Time values should decrease along trajectory."""
        times = []
        for x, t in diffusion_sampler.sample_with_trajectory(shape=(1, 1, 8, 8), yield_every=5):
            times.append(t)
        # First time should be 1.0, times should generally decrease
        assert times[0] == 1.0
        assert times[-1] < times[0]


class TestDiffusionSamplerODE:
    """Tests for sample_ode method (deterministic sampling)."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_sample_ode_returns_tensor(self, diffusion_sampler):
        """sample_ode should return a tensor."""
        samples = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8))
        assert isinstance(samples, torch.Tensor)

    def test_sample_ode_correct_shape(self, diffusion_sampler):
        """sample_ode should return correct shape."""
        samples = diffusion_sampler.sample_ode(shape=(3, 1, 8, 8))
        assert samples.shape == (3, 1, 8, 8)

    def test_sample_ode_different_methods(self, diffusion_sampler):
        """sample_ode should support different ODE methods."""
        for method in ["euler", "heun", "rk4"]:
            samples = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method=method)
            assert samples.shape == (2, 1, 8, 8)


class TestDiffusionSamplerPredictorCorrector:
    """Tests for sample_predictor_corrector method."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_predictor_corrector_returns_tensor(self, diffusion_sampler):
        """sample_predictor_corrector should return a tensor."""
        samples = diffusion_sampler.sample_predictor_corrector(
            shape=(2, 1, 8, 8), n_corrector_steps=2
        )
        assert isinstance(samples, torch.Tensor)

    def test_predictor_corrector_correct_shape(self, diffusion_sampler):
        """sample_predictor_corrector should return correct shape."""
        samples = diffusion_sampler.sample_predictor_corrector(
            shape=(3, 1, 8, 8), n_corrector_steps=1
        )
        assert samples.shape == (3, 1, 8, 8)


class TestDiffusionSamplerStatistics:
    """Tests for compute_sample_statistics method."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_statistics_returns_dict(self, diffusion_sampler):
        """compute_sample_statistics should return a dictionary."""
        samples = torch.randn(5, 1, 8, 8)
        stats = diffusion_sampler.compute_sample_statistics(samples)
        assert isinstance(stats, dict)

    def test_statistics_contains_expected_keys(self, diffusion_sampler):
        """Statistics should contain expected keys."""
        samples = torch.randn(5, 1, 8, 8)
        stats = diffusion_sampler.compute_sample_statistics(samples)
        # Check for some expected keys
        assert "mean" in stats or "sample_mean" in stats


class TestDiffusionSamplerTemperatureScaling:
    """Tests for temperature scaling feature."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_training_temperature_property(self, diffusion_sampler):
        """Training temperature should be settable and gettable."""
        assert diffusion_sampler.training_temperature is None
        diffusion_sampler.training_temperature = 2.27
        assert diffusion_sampler.training_temperature == 2.27

    def test_compute_temperature_scale_linear(self, diffusion_sampler):
        """Linear scaling should compute T_target / T_train."""
        diffusion_sampler.training_temperature = 2.0
        scale = diffusion_sampler.compute_temperature_scale(4.0, method="linear")
        assert scale == 2.0

    def test_compute_temperature_scale_sqrt(self, diffusion_sampler):
        """Sqrt scaling should compute sqrt(T_target / T_train)."""
        diffusion_sampler.training_temperature = 4.0
        scale = diffusion_sampler.compute_temperature_scale(4.0, method="sqrt")
        assert scale == 1.0

    def test_compute_temperature_scale_no_training_temp(self, diffusion_sampler):
        """Without training temp, should return target temperature."""
        scale = diffusion_sampler.compute_temperature_scale(2.5, method="linear")
        assert scale == 2.5

    def test_sample_at_temperature(self, diffusion_sampler):
        """sample_at_temperature should return samples with correct shape."""
        diffusion_sampler.training_temperature = 2.27
        samples = diffusion_sampler.sample_at_temperature(
            shape=(2, 1, 8, 8),
            target_temperature=3.0,
        )
        assert samples.shape == (2, 1, 8, 8)


class TestDiffusionSamplerTrajectoryWithMetadata:
    """Tests for sample_trajectory_with_metadata method."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=20,
        )

    def test_metadata_trajectory_is_generator(self, diffusion_sampler):
        """sample_trajectory_with_metadata should return a generator."""
        from types import GeneratorType
        trajectory = diffusion_sampler.sample_trajectory_with_metadata(
            shape=(1, 1, 8, 8), n_frames=5
        )
        assert isinstance(trajectory, GeneratorType)

    def test_metadata_trajectory_yields_dicts(self, diffusion_sampler):
        """Trajectory should yield dictionaries with metadata."""
        for frame in diffusion_sampler.sample_trajectory_with_metadata(
            shape=(1, 1, 8, 8), n_frames=3
        ):
            assert isinstance(frame, dict)
            assert "x" in frame
            assert "t" in frame
            assert "step" in frame
            assert "progress" in frame
            assert "sigma" in frame
            break

    def test_metadata_trajectory_frame_spacing_linear(self, diffusion_sampler):
        """Linear spacing should yield evenly spaced frames."""
        frames = list(diffusion_sampler.sample_trajectory_with_metadata(
            shape=(1, 1, 8, 8), n_frames=5, frame_spacing="linear"
        ))
        assert len(frames) > 0
        # Progress should increase monotonically
        progress_values = [f["progress"] for f in frames]
        assert progress_values[0] <= progress_values[-1]

    def test_metadata_trajectory_frame_spacing_log(self, diffusion_sampler):
        """Log spacing should yield more frames near the end."""
        frames = list(diffusion_sampler.sample_trajectory_with_metadata(
            shape=(1, 1, 8, 8), n_frames=5, frame_spacing="log"
        ))
        assert len(frames) > 0

    def test_metadata_trajectory_frame_spacing_cosine(self, diffusion_sampler):
        """Cosine spacing should yield smooth acceleration."""
        frames = list(diffusion_sampler.sample_trajectory_with_metadata(
            shape=(1, 1, 8, 8), n_frames=5, frame_spacing="cosine"
        ))
        assert len(frames) > 0


class TestDiffusionSamplerGenerateBatch:
    """Tests for generate_batch method."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_generate_batch_returns_correct_count(self, diffusion_sampler):
        """generate_batch should return requested number of samples."""
        samples = diffusion_sampler.generate_batch(
            n_samples=5, lattice_size=8, batch_size=2
        )
        assert samples.shape[0] == 5

    def test_generate_batch_correct_shape(self, diffusion_sampler):
        """generate_batch should return correct shape (no channel dim)."""
        samples = diffusion_sampler.generate_batch(
            n_samples=4, lattice_size=16, batch_size=2
        )
        assert samples.shape == (4, 16, 16)

    def test_generate_batch_discrete_by_default(self, diffusion_sampler):
        """generate_batch should discretize by default."""
        samples = diffusion_sampler.generate_batch(
            n_samples=3, lattice_size=8, discretize=True
        )
        unique = torch.unique(samples)
        assert all(v in [-1, 1] for v in unique.tolist())

    def test_generate_batch_different_methods(self, diffusion_sampler):
        """generate_batch should support different sampling methods."""
        for method in ["sde", "ode", "pc"]:
            samples = diffusion_sampler.generate_batch(
                n_samples=2, lattice_size=8, sampling_method=method
            )
            assert samples.shape == (2, 8, 8)


class TestDiffusionSamplerODEReproducibility:
    """Tests for deterministic ODE sampling reproducibility."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_ode_euler_reproducibility(self, diffusion_sampler):
        """ODE euler method should produce identical results with same seed."""
        torch.manual_seed(42)
        samples1 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="euler")

        torch.manual_seed(42)
        samples2 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="euler")

        assert torch.allclose(samples1, samples2, atol=1e-6)

    def test_ode_heun_reproducibility(self, diffusion_sampler):
        """ODE heun method should produce identical results with same seed."""
        torch.manual_seed(123)
        samples1 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="heun")

        torch.manual_seed(123)
        samples2 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="heun")

        assert torch.allclose(samples1, samples2, atol=1e-6)

    def test_ode_rk4_reproducibility(self, diffusion_sampler):
        """ODE rk4 method should produce identical results with same seed."""
        torch.manual_seed(456)
        samples1 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="rk4")

        torch.manual_seed(456)
        samples2 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="rk4")

        assert torch.allclose(samples1, samples2, atol=1e-6)

    def test_ode_different_seeds_produce_different_samples(self, diffusion_sampler):
        """Different seeds should produce different ODE samples."""
        torch.manual_seed(1)
        samples1 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="euler")

        torch.manual_seed(2)
        samples2 = diffusion_sampler.sample_ode(shape=(2, 1, 8, 8), method="euler")

        # Should not be identical
        assert not torch.allclose(samples1, samples2, atol=1e-6)


class TestDiffusionSamplerCheckpoint:
    """Tests for checkpoint save/load functionality."""

    @pytest.fixture
    def diffusion_sampler(self):
        """Create a diffusion sampler for testing."""
        model = ScoreNetwork(in_channels=1, base_channels=16, time_embed_dim=32, num_blocks=2)
        return DiffusionSampler(
            score_network=model,
            num_steps=10,
        )

    def test_save_checkpoint(self, diffusion_sampler, tmp_path):
        """save_checkpoint should create a valid checkpoint file."""
        checkpoint_path = tmp_path / "test_checkpoint.pt"
        diffusion_sampler.save_checkpoint(str(checkpoint_path))
        assert checkpoint_path.exists()

    def test_save_checkpoint_contains_required_keys(self, diffusion_sampler, tmp_path):
        """Checkpoint should contain model state dict and config."""
        checkpoint_path = tmp_path / "test_checkpoint.pt"
        diffusion_sampler.save_checkpoint(str(checkpoint_path))

        checkpoint = torch.load(checkpoint_path, weights_only=False)
        assert "model_state_dict" in checkpoint
        assert "num_steps" in checkpoint
        assert "diffusion_config" in checkpoint

    def test_save_checkpoint_with_model_config(self, diffusion_sampler, tmp_path):
        """Checkpoint should save model config when provided."""
        checkpoint_path = tmp_path / "test_checkpoint.pt"
        config = {"in_channels": 1, "base_channels": 16}
        diffusion_sampler.save_checkpoint(str(checkpoint_path), model_config=config)

        checkpoint = torch.load(checkpoint_path, weights_only=False)
        assert "model_config" in checkpoint
        assert checkpoint["model_config"] == config

    def test_load_from_checkpoint(self, diffusion_sampler, tmp_path):
        """from_checkpoint should create a working sampler."""
        checkpoint_path = tmp_path / "test_checkpoint.pt"
        config = {"in_channels": 1, "base_channels": 16, "time_embed_dim": 32, "num_blocks": 2}
        diffusion_sampler.save_checkpoint(str(checkpoint_path), model_config=config)

        loaded_sampler = DiffusionSampler.from_checkpoint(str(checkpoint_path), num_steps=10)
        assert loaded_sampler is not None
        assert loaded_sampler.num_steps == 10

    def test_load_checkpoint_samples_match_shape(self, diffusion_sampler, tmp_path):
        """Loaded sampler should produce correct output shapes."""
        checkpoint_path = tmp_path / "test_checkpoint.pt"
        config = {"in_channels": 1, "base_channels": 16, "time_embed_dim": 32, "num_blocks": 2}
        diffusion_sampler.save_checkpoint(str(checkpoint_path), model_config=config)

        loaded_sampler = DiffusionSampler.from_checkpoint(str(checkpoint_path), num_steps=10)
        samples = loaded_sampler.sample(shape=(2, 1, 8, 8))
        assert samples.shape == (2, 1, 8, 8)


class TestPretrainedDiffusionSampler:
    """Tests for PretrainedDiffusionSampler (heuristic mode)."""

    def test_init_creates_sampler(self):
        """PretrainedDiffusionSampler should initialize correctly."""
        sampler = PretrainedDiffusionSampler(lattice_size=16, num_steps=50)
        assert sampler.lattice_size == 16
        assert sampler.num_steps == 50

    def test_sample_heuristic_returns_tensor(self):
        """sample_heuristic should return a tensor."""
        sampler = PretrainedDiffusionSampler(lattice_size=8, num_steps=20)
        samples = sampler.sample_heuristic(batch_size=2, temperature=2.27)
        assert isinstance(samples, torch.Tensor)

    def test_sample_heuristic_correct_shape(self):
        """sample_heuristic should return correct shape."""
        sampler = PretrainedDiffusionSampler(lattice_size=16, num_steps=20)
        samples = sampler.sample_heuristic(batch_size=3, temperature=2.27)
        assert samples.shape == (3, 1, 16, 16)

    def test_sample_heuristic_discrete_values(self):
        """sample_heuristic should produce discrete ±1 values."""
        sampler = PretrainedDiffusionSampler(lattice_size=8, num_steps=20)
        samples = sampler.sample_heuristic(batch_size=5, temperature=2.27)
        samples_2d = samples.squeeze(1)
        unique = torch.unique(samples_2d)
        assert all(v in [-1, 1] for v in unique.tolist())

    def test_sample_with_trajectory_generator(self):
        """sample_with_trajectory should return a generator."""
        from types import GeneratorType
        sampler = PretrainedDiffusionSampler(lattice_size=8, num_steps=20)
        trajectory = sampler.sample_with_trajectory(shape=(1, 1, 8, 8))
        assert isinstance(trajectory, GeneratorType)

    def test_sample_with_trajectory_yields_tuples(self):
        """sample_with_trajectory should yield (tensor, time) tuples."""
        sampler = PretrainedDiffusionSampler(lattice_size=8, num_steps=20)
        frames = list(sampler.sample_with_trajectory(shape=(1, 1, 8, 8), yield_every=5))
        assert len(frames) > 0
        for x, t in frames:
            assert isinstance(x, torch.Tensor)
            assert isinstance(t, float)
