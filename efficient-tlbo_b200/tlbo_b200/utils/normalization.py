"""y-normalisation helpers (reference tlbo/utils/normalization.py:4-28)."""
import numpy as np


def zero_one_normalization(X, lower=None, upper=None):
    lower = np.min(X, axis=0) if lower is None else lower
    upper = np.max(X, axis=0) if upper is None else upper
    return np.true_divide(X - lower, upper - lower), lower, upper


def zero_one_unnormalization(X_normalized, lower, upper):
    return lower + (upper - lower) * X_normalized


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    mean = np.mean(X, axis=0) if mean is None else mean
    std = np.std(X, axis=0) if std is None else std
    return (X - mean) / std, mean, std


def zero_mean_unit_var_unnormalization(X_normalized, mean, std):
    return X_normalized * std + mean
