"""Host-side mirror of the reference's `torch_utils` package for the render/StyleGAN2-op hot path."""
