"""Optimizer wrapper that drives the fault-tolerance protocol
(reference: /root/reference/torchft/optim.py:24-63).

``zero_grad()`` starts the step's quorum, ``step()`` only applies the update when
the replica group agrees the step is clean (``Manager.should_commit``).
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any, Dict, List, Mapping, Optional

import torch
from torch.optim import Optimizer

if TYPE_CHECKING:
    from torchft_b200.manager import Manager


class OptimizerWrapper(Optimizer):
    """Wrap any optimizer (torch or :class:`~torchft_b200.ops.fused.FlatAdamW`).

        optim = OptimizerWrapper(manager, torch.optim.AdamW(m.parameters()))
        optim.zero_grad()        # -> manager.start_quorum()
        loss.backward()
        optim.step()             # -> if manager.should_commit(): inner.step()
    """

    def __init__(self, manager: "Manager", optim: Any) -> None:
        # deliberately no Optimizer.__init__: all state lives in the wrapped optimizer
        self.optim = optim
        self.manager = manager

    def add_param_group(self, param_group: Dict[str, Any]) -> None:
        self.optim.add_param_group(param_group)

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        self.optim.load_state_dict(state_dict)

    def state_dict(self) -> Dict[str, Any]:
        return self.optim.state_dict()

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.manager.start_quorum()
        self.optim.zero_grad(set_to_none)

    def step(self, closure: Optional[object] = None) -> None:
        assert closure is None, "optimizers that use closures are not supported"
        if self.manager.should_commit():
            self.optim.step()

    @property
    def param_groups(self) -> List[Dict[str, Any]]:  # type: ignore[override]
        return self.optim.param_groups

    @property
    def state(self) -> Mapping[torch.Tensor, Any]:  # type: ignore[override]
        return self.optim.state
