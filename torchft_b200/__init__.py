"""torchft_b200 -- per-step fault-tolerant training, built for 8xB200 / NVLink 5 / NVSwitch.

Same capabilities and API surface as meta-pytorch/torchft (Lighthouse quorum
server, Manager, reconfigurable process groups, fault-tolerant DDP/HSDP,
LocalSGD/DiLoCo, live checkpoint recovery), with the control plane in C++ and
the data plane as hand-written sm_100a kernels over NVLink peer memory.
"""

import logging as _logging

from torchft_b200.data import DistributedSampler
from torchft_b200.ddp import DistributedDataParallel, FlatDistributedDataParallel, PureDistributedDataParallel
from torchft_b200.manager import Manager, WorldSizeMode
from torchft_b200.optim import OptimizerWrapper as Optimizer
from torchft_b200.otel import setup_logger as _setup_logger
from torchft_b200.process_group import (
    ManagedProcessGroup,
    ProcessGroupDummy,
    ProcessGroupGloo,
    ProcessGroupNCCL,
    ProcessGroupXCCL,
)
from torchft_b200.baby import ProcessGroupBabyGloo, ProcessGroupBabyNCCL, ProcessGroupBabyXCCL

for _name in ("torchft_quorums", "torchft_commits", "torchft_errors"):
    _setup_logger(_name)


def __getattr__(name: str):  # lazy: needs CUDA at construction time only
    if name == "ProcessGroupB200":
        from torchft_b200.parallel.process_group_b200 import ProcessGroupB200

        return ProcessGroupB200
    raise AttributeError(name)


__all__ = [
    "DistributedDataParallel",
    "FlatDistributedDataParallel",
    "PureDistributedDataParallel",
    "DistributedSampler",
    "Manager",
    "WorldSizeMode",
    "Optimizer",
    "ManagedProcessGroup",
    "ProcessGroupB200",
    "ProcessGroupNCCL",
    "ProcessGroupXCCL",
    "ProcessGroupBabyNCCL",
    "ProcessGroupBabyXCCL",
    "ProcessGroupBabyGloo",
    "ProcessGroupGloo",
    "ProcessGroupDummy",
]
