"""Tri-plane volumetric renderer (mirror of the reference's training/volumetric_rendering)."""
