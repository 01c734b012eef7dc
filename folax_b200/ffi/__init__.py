"""jax.ffi side of the drop-in boundary: xla_ffi_shim.cc (handlers) and jax_binding.py (loss classes); neither can run in an image without JAX."""
