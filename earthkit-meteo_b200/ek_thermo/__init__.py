"""ek_thermo -- B200-native (sm_100a) kernels for the earthkit-meteo thermo hot path.

    from ek_thermo import thermo          # drop-in for `from earthkit.meteo import thermo` on CUDA tensors
    theta = thermo.potential_temperature(t, p)

    from ek_thermo import fused            # one pass, many outputs
    fields = fused.suite_tqp(t, q, p)      # {"theta", "es", "rh", "td", "tv"}

    from ek_thermo import host             # numpy in, numpy out, computed on the GPU (chunked H2D / kernel / D2H)
    theta = host.thermo.potential_temperature(t_np, p_np)

Importing this package loads libek_thermo.so and fails loudly if it has not been built: there is no
CPU or eager-PyTorch fallback anywhere.
"""
from . import _backend, fused, host, hostpipe, partition, thermo, vertical, wind  # noqa: F401
from ._backend import EkThermoError, launch_count, set_launch_config, version  # noqa: F401

__all__ = ["thermo", "vertical", "wind", "fused", "partition", "hostpipe", "host", "version", "launch_count", "set_launch_config", "EkThermoError"]
