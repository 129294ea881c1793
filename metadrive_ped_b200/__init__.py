"""metadrive_ped_b200 — a B200-native batched MetaDrive step behind the reference's Gymnasium surface.

Only the step path lives here (include/mdstep.h is the boundary): hand-written sm_100a kernels in csrc/, the ctypes
binding (lib, sim), host-side scene tables (scene, library) and the drop-in env classes (envs)."""
from .envs import (BatchedMetaDriveEnv, BatchedMultiAgentEnv, MetaDriveEnv, MultiAgentBottleneckEnv,  # noqa: F401
                   MultiAgentIntersectionEnv, MultiAgentMetaDrive, MultiAgentParkingLotEnv, MultiAgentRoundaboutEnv,
                   MultiAgentTollgateEnv,
                   SafeMetaDriveEnv, TopDownMetaDrive, TopDownSingleFrameMetaDriveEnv)

__version__ = "0.1.0"
