"""Bindings of the C ABI for callers that do not hold torch CUDA tensors."""
