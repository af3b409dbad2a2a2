from .gae import compute_gae, standardize_

__all__ = ["compute_gae", "standardize_"]
