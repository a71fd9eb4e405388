"""Optimizer facade that drives the per-step fault-tolerance protocol.

Behavioural parity with the reference's ``torchft/optim.py:24-63``: ``zero_grad()`` opens the
step (starts the quorum, which overlaps the forward pass when the manager uses an async quorum)
and ``step()`` applies the update only if every rank of the replica group — and therefore every
participating replica — voted the step clean (``Manager.should_commit``).

Implementation notes (ours): the facade owns no optimizer state. Everything that is not part of
the protocol is *forwarded* to the wrapped object — by a generated method for names that
``torch.optim.Optimizer`` defines itself, by ``__getattr__`` for the rest — so it works for any
``torch.optim`` optimizer and for duck-typed ones such as :class:`torchft_b200.ops.fused.FlatAdamW`,
and LR schedulers (which require an ``Optimizer`` instance and read ``param_groups``) keep working.
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any, Callable, Optional

from torch.optim import Optimizer

if TYPE_CHECKING:
    from torchft_b200.manager import Manager

__all__ = ["OptimizerWrapper"]

_DATA = ("param_groups", "state", "defaults")

# Optimizer methods that must reach the wrapped optimizer instead of the (uninitialised) base class
_FORWARDED = ("add_param_group", "load_state_dict", "state_dict", "register_step_pre_hook", "register_step_post_hook",
              "register_state_dict_pre_hook", "register_state_dict_post_hook", "register_load_state_dict_pre_hook",
              "register_load_state_dict_post_hook")


class OptimizerWrapper(Optimizer):
    """``OptimizerWrapper(manager, inner)``; exported as ``torchft_b200.Optimizer``.

        optim = OptimizerWrapper(manager, torch.optim.AdamW(model.parameters()))
        optim.zero_grad()        # manager.start_quorum()
        loss.backward()          # gradients averaged over the live replicas
        optim.step()             # inner.step() iff manager.should_commit()

    ``last_step_committed`` tells the training loop whether the most recent ``step()`` was applied.
    """

    def __init__(self, manager: "Manager", optim: Any) -> None:
        # no Optimizer.__init__ on purpose: defaults / param_groups / state all belong to `optim`
        self.__dict__["optim"] = optim
        self.__dict__["manager"] = manager
        self.__dict__["last_step_committed"] = False

    # -- the protocol --------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:  # type: ignore[override]
        self.manager.start_quorum()
        self.optim.zero_grad(set_to_none)

    def step(self, closure: Optional[Callable[[], float]] = None) -> None:  # type: ignore[override]
        if closure is not None:
            raise NotImplementedError("closure-based optimizers re-evaluate the loss outside the commit protocol")
        self.last_step_committed = bool(self.manager.should_commit())
        if self.last_step_committed:
            self.optim.step()

    # -- everything else belongs to the wrapped optimizer ------------------------
    def __getattr__(self, name: str) -> Any:  # only called when normal lookup fails
        return getattr(self.__dict__["optim"], name)

    def __setattr__(self, name: str, value: Any) -> None:
        # the optimizer's DATA lives in the wrapped object; anything else (e.g. the `step` patch an LR
        # scheduler installs to count calls) stays on the facade, exactly as for a plain Optimizer
        if name in _DATA:
            setattr(self.__dict__["optim"], name, value)
        else:
            self.__dict__[name] = value

    def __repr__(self) -> str:
        return f"OptimizerWrapper({self.optim!r})"


def _forward(name: str) -> Callable[..., Any]:
    def method(self: OptimizerWrapper, *args: Any, **kwargs: Any) -> Any:
        return getattr(self.optim, name)(*args, **kwargs)

    method.__name__ = name
    method.__doc__ = f"Forwarded to the wrapped optimizer's ``{name}``."
    return method


for _name in _FORWARDED:
    if hasattr(Optimizer, _name):
        setattr(OptimizerWrapper, _name, _forward(_name))
del _name
