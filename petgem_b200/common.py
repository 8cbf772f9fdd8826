"""Parameter file, printing and timers: the glue around the hot path.

Mirrors the surface of ``petgem/common.py``: ``InputParameters`` reads the same
YAML schema (``model/run/output``, common.py:129-361) with the same defaults
(``run.cuda`` False, ``output.vtk`` False, scratch directory) and the same failure
convention (message on the master rank, ``exit(-1)``); ``Print.master`` and
``Timers`` keep their call signatures.  No colour codes, no singletons.
"""
from __future__ import annotations

import os
import sys
import time



def _rank() -> int:
    for key in ("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


class Print(object):
    """common.py:30-124."""

    def __init__(self, text, color_code=None):
        print(text)
        sys.stdout.flush()

    @classmethod
    def master(cls, text, color_code=None):
        if _rank() == 0:
            print(text)
            sys.stdout.flush()

    @classmethod
    def header(cls):
        if _rank() == 0:
            bar = "%" * 75
            for line in (bar, "%%%" + "PETGEM hot path -- B200 native".center(69) + "%%%", bar):
                print(line)
            sys.stdout.flush()


def _fail(msg: str):
    Print.master("     " + msg)
    exit(-1)


class InputParameters(object):
    """common.py:129-361: YAML -> .model/.run/.output dictionaries with the reference's checks."""

    _CONSISTENCY = " Please, verify the parameter file consistency."

    def __init__(self, params, parEnv=None):
        import yaml

        with open(params, "r") as f:
            inputs = yaml.safe_load(f)
        self.model, self.run, self.output = inputs["model"], inputs["run"], inputs["output"]
        mode = self.model.get("mode")
        if mode is None:
            _fail("Modeling mode not provided." + self._CONSISTENCY)
        if mode not in ("csem", "mt"):
            _fail("Modeling mode not supported.")
        if mode not in self.model:
            _fail(mode + " parameters not provided." + self._CONSISTENCY)
        from_file, npol = self._verify(mode, self.model[mode])
        self.run.update({"conductivity_from_file": from_file, "num_polarizations": npol})
        for key in ("mesh", "receivers"):
            if key not in self.model:
                _fail(key + " parameter not provided." + self._CONSISTENCY)
        if "nord" not in self.run:
            _fail("nord parameter not provided." + self._CONSISTENCY)
        if self.run["nord"] < 1 or self.run["nord"] > 6:
            _fail("Vector finite element basis order not supported. Please, select a valid order (1,2,3,4,5,6).")
        if "cuda" not in self.run:
            self.run["cuda"] = False
        elif self.run["cuda"] is not True and self.run["cuda"] is not False:
            _fail("cuda option not supported. Please, select a valid order (True/False).")
        self.output.setdefault("vtk", False)
        if "directory" not in self.output:
            _fail("output directory parameter not provided." + self._CONSISTENCY)
        rank = parEnv.rank if parEnv is not None else _rank()
        if rank == 0:
            os.makedirs(self.output["directory"], exist_ok=True)
        if "directory_scratch" not in self.output:
            self.output.update({"directory_scratch": self.output["directory"], "remove_scratch": False})
        else:
            if rank == 0:
                os.makedirs(self.output["directory_scratch"], exist_ok=True)
            # common.py:343-352 always removes a user-given scratch directory after the run; an explicit
            # `remove_scratch: False` in the output section keeps it (x{i}.dat stay readable)
            self.output["remove_scratch"] = self.output.get("remove_scratch", True) is not False

    def _verify(self, mode, data):
        sigma = data.get("sigma")
        if sigma is None:
            _fail(mode + " parameters not provided." + self._CONSISTENCY)
        has_file = "file" in sigma
        has_arrays = "horizontal" in sigma and "vertical" in sigma
        if has_file == has_arrays or (has_file and ("horizontal" in sigma or "vertical" in sigma)):
            _fail("sigma parameters invalid." + self._CONSISTENCY)
        if mode == "csem":
            src = data.get("source")
            if src is None:
                _fail("source parameters not provided." + self._CONSISTENCY)
            names = ["frequency", "position", "azimuth", "dip", "current", "length"]
            if len(src) != 6:
                _fail("number of source parameters is not consistent." + self._CONSISTENCY)
            for n in names:
                if n not in src:
                    _fail(n + " parameter not provided." + self._CONSISTENCY)
            return has_file, int(1)
        for n in ("frequency", "polarization"):
            if n not in data:
                _fail(n + " parameter not provided for model." + self._CONSISTENCY)
        return has_file, len(data["polarization"])


class Timer:
    """common.py:367-399."""

    def __init__(self, elapsed=0.0):
        self._start = None
        self._elapsed = elapsed

    def start(self):
        self._start = time.time()

    def stop(self):
        if self._start is not None:
            self._elapsed += time.time() - self._start
            self._start = None

    def reset(self):
        self._elapsed = 0.0

    @property
    def elapsed(self):
        return self._elapsed


class Timers:
    """common.py:402-473: named timers, ``Timers()["Assembly"].start()``; one registry per process."""

    _registry = {}

    def __init__(self, opath=None):
        self._opath = opath

    def __getitem__(self, key):
        return Timers._registry.setdefault(key, Timer())

    def elapsed(self, key):
        return Timers._registry[key].elapsed if key in Timers._registry else 0.0

    def items(self):
        return {k: v.elapsed for k, v in Timers._registry.items()}


def unitary_test():
    """Unitary test for common.py script."""


__all__ = ["Print", "InputParameters", "Timer", "Timers"]
