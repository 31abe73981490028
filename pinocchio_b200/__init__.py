"""pinocchio_b200 — B200-native batched rigid-body dynamics behind Pinocchio's *InParallel interface.

Only the many-configuration hot path of stack-of-tasks/pinocchio is rebuilt here:
RNEA / ABA / CRBA / computeRNEADerivatives / computeABADerivatives over batches whose columns are
configurations.  The compute path is hand-written sm_100a CUDA behind a C ABI
(include/pinocchio_b200.h); this package is the host-side mirror of the reference interface.
"""
from .model import (JOINT_FREEFLYER, JOINT_PLANAR, JOINT_PRISMATIC_UNALIGNED, JOINT_PX, JOINT_PY, JOINT_PZ,  # noqa: F401
                    JOINT_REVOLUTE_UNALIGNED, JOINT_RX, JOINT_RY, JOINT_RZ, JOINT_SPHERICAL, SE3, Inertia, Model, buildModelFromUrdf, buildSampleModelHumanoid,
                    buildSampleModelHumanoidRandom, buildSampleModelManipulator)
from .joint_configuration import (LibcRand, batched_random_configuration, batched_random_tangent, integrate,  # noqa: F401
                                  neutral, randomConfiguration)
from .pool import (ModelPool, abaEulerStepInParallel, abaInParallel, computeABADerivativesInParallel,  # noqa: F401
                   computeGeneralizedGravityInParallel, computeMinverseInParallel, computeRNEADerivativesInParallel,
                   crbaInParallel, crbaPackedInParallel, expandPackedCrba, integrateInParallel, nonLinearEffectsInParallel, pin_host, rneaInParallel, unpin_host)

__version__ = "0.1.0"
