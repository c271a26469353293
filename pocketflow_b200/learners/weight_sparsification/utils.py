"""Utility functions for the weight sparsification learner
(/root/reference/learners/weight_sparsification/utils.py:19-39)."""


def get_maskable_vars(trainable_vars):
    """Kernels of conv2d / dense layers and slim pointwise / final 1x1 conv weights."""
    vars_kernel = [var for var in trainable_vars if 'kernel' in var.name]
    vars_ptconv = [var for var in trainable_vars if 'pointwise/weights' in var.name]
    vars_fnconv = [var for var in trainable_vars if 'Conv2d_1c_1x1/weights' in var.name]
    return vars_kernel + vars_ptconv + vars_fnconv
