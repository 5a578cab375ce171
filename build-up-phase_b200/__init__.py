"""B200-native ray-tracing core: host-side Python mirror of the C ABI in include/rtcore.h.

`scenes` is pure numpy (scene descriptions); `rtcore` binds librtcore.so (CUDA, sm_100a) and raises
loudly when the library or a GPU is missing — there is no CPU fallback in the product.
"""
from . import scenes  # noqa: F401

__all__ = ["scenes", "rtcore", "build"]
