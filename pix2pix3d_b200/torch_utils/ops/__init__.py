"""Functional op surface of the reference (`torch_utils.ops.*`) backed by libp3d.so on CUDA tensors."""
