"""Estimator contract shared with the reference (probaforms/models/interfaces.py:6-43).

``GenModel`` is an ``nn.Module`` with an sklearn-style ``fit(X, C)`` /
``sample(C)`` pair; the reference's test-suite discovers models through
``GenModel.__subclasses__()`` (tests/test_models.py:6-10), so the class keeps
that name and base.
"""
import torch.nn as nn


class GenModel(nn.Module):
    """Conditional generative model: ``fit(X, C)`` learns p(x|c), ``sample(C)`` draws from it."""

    def fit(self, X, C):
        """X: [n, var_size] array, C: [n, cond_size] array or None."""
        raise NotImplementedError

    def sample(self, C):
        """C: [n, cond_size] array of conditions, or an int number of rows to draw."""
        raise NotImplementedError
