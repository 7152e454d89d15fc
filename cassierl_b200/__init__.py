"""cassierl_b200 -- B200-native batched replacement for CassieRL/cassierl's libcassie2d step path.

The product path is the CUDA library `lib/libcassie2d.so` (C-ABI in include/cassie2d.h); this
package is the host-side mirror of the reference's Python interface
(rllab/envs/cassie2d.py, cassie_stand2d.py, cassie2d_structs.py) over a batch dimension.
There is no CPU fallback: importing `cassierl_b200.lib` fails loudly if the library is missing,
and every call fails if no CUDA device is usable.
"""
from .structs import (ControllerForce, ControllerOsc, ControllerPd, ControllerTorque,  # noqa: F401
                      InterfaceStructConverter, StateGeneral, StateOperationalSpace)

__all__ = ["ControllerForce", "ControllerOsc", "ControllerPd", "ControllerTorque", "StateGeneral",
           "StateOperationalSpace", "InterfaceStructConverter"]
