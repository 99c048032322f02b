"""Hyper-parameter containers with the behaviour MLatom callers rely on
(reference: mlatom/model_cls.py:856-958): ``model.hyperparameters['sigma'].value``,
``.minval/.maxval/.optimization_space``, dict-style update with plain numbers, attribute access."""
from __future__ import annotations

from collections import UserDict
from typing import Any, Iterable, Optional

import numpy as np


class hyperparameter:
    def __init__(self, value: Any = None, optimization_space: str = "linear", dtype=None, name: str = "",
                 minval: Any = None, maxval: Any = None, step: Any = None, choices: Iterable[Any] = (), **_ignored):
        self.name = name
        self.dtype = dtype if dtype else (None if value is None else type(value))
        self.value = value
        self.optimization_space = optimization_space
        self.minval, self.maxval, self.step = minval, maxval, step
        self.choices = list(choices)

    def __setattr__(self, key, value):
        if key == "value" and getattr(self, "dtype", None) is not None and value is not None:
            caster = np.array if self.dtype is np.ndarray else self.dtype
            if not isinstance(value, self.dtype):
                value = caster(value)
        object.__setattr__(self, key, value)

    def update(self, other: "hyperparameter") -> None:
        self.__dict__.update(other.__dict__)

    def copy(self) -> "hyperparameter":
        return hyperparameter(**self.__dict__)

    def __repr__(self):
        return f"hyperparameter {self.__dict__}"


class hyperparameters(UserDict):
    """Mapping name -> hyperparameter; assigning a plain value updates ``.value`` in place."""

    def __setitem__(self, key, value):
        if isinstance(value, hyperparameter):
            if key in self.data:
                self.data[key].update(value)
            else:
                self.data[key] = value
        elif key in self.data:
            self.data[key].value = value
        else:
            self.data[key] = hyperparameter(value=value, name=key)

    def __getattr__(self, key):
        data = self.__dict__.get("data", {})
        if key in data:
            return data[key].value
        raise AttributeError(key)

    def __setattr__(self, key, value):
        if key == "data" or key.startswith("__") or key in self.__dict__:
            object.__setattr__(self, key, value)
        else:
            self[key] = value

    def copy(self, keys: Optional[Iterable[str]] = None) -> "hyperparameters":
        keys = list(self.keys()) if keys is None else list(keys)
        return hyperparameters({k: self.data[k].copy() for k in keys})
