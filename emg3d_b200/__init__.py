"""emg3d_b200 -- the multigrid hot path of emg3d, rebuilt for NVIDIA B200.

Python host code calling hand-written sm_100a CUDA kernels through a C ABI
(``libemg3d_b200.so``, ``include/emg3d_b200.h``); same solver API as
``emg3d.solver``.  There is no CPU fallback.
"""
from emg3d_b200._lib import Emg3dB200Error
from emg3d_b200.meshes import BaseMesh, TensorMesh
from emg3d_b200.fields import (Field, SourceField, DeviceField, get_source_field, get_magnetic_field,
                               get_receiver)
from emg3d_b200.models import Model, VolumeModel
from emg3d_b200 import batch, core, maps, solver
from emg3d_b200.batch import solve_many
from emg3d_b200.solver import solve, solve_source, Workspace, __version__

__all__ = ['solve', 'solve_source', 'Model', 'VolumeModel', 'Field',
           'get_source_field', 'get_magnetic_field', 'get_receiver', 'SourceField', 'DeviceField', 'maps',
           'solve_many', 'batch', 'TensorMesh', 'BaseMesh', 'core', 'solver', 'Workspace',
           'Emg3dB200Error']
