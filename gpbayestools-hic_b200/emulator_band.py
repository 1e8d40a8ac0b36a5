"""Drop-in for the reference's `EmulatorBAND` (src/emulator_BAND.py): the wrapper around the BAND
collaboration's surmise emulators (PCGP / PCSK / PCGPwImpute / PCGPwM).

Training stays what it is in the reference -- a call into surmise, offline, on the CPU
(src/emulator_BAND.py:258-292) -- and needs the surmise package.  `predict` does not: the trained
surmise object's fit information (`emu._info`: theta, pct, scale, offset, extravar, emulist[k] =
{hypcov, hypind, nug, Vh, pw, sig2}) is copied to the GPU once and every call runs kernels (a)+(b)
(csrc/pc_predict.cuh KIND 2).  surmise 0.2.1 is neither installed in the build image nor part of
the reference tree, so this path is checked against a restatement of its published predict
algorithm only: PARITY UNPINNED against surmise itself."""
from __future__ import annotations

import logging

import numpy as np

from . import parse_model_parameter_file
from .device import DeviceEmulator
from .emulator import Emulator, read_training_pickle
from .state import EmulatorState

log = logging.getLogger(__name__)

METHODS = ("PCGP", "PCSK", "PCGPwImpute", "PCGPwM")


class EmulatorBAND(Emulator):
    def __init__(self, training_set_path=".", parameter_file="ABCD.txt", method="PCGP", logTrafo=False,
                 parameterTrafoPCA=False, max_rel_uncertainty_data=0.1, exp_and_cov_diagonal=False):
        if exp_and_cov_diagonal and not logTrafo:
            raise ValueError("exp_and_cov_diagonal can only be set to True if logTrafo is True.")
        if method not in METHODS:
            raise ValueError("Requested method not implemented!")
        self.method_ = method
        self.logTrafo_ = logTrafo
        self.parameterTrafoPCA_ = bool(parameterTrafoPCA)
        self.max_rel_uncertainty_data_ = max_rel_uncertainty_data
        self.exp_and_cov_diagonal_ = exp_and_cov_diagonal
        self.perform_no_PCA_ = False
        self.design_points, self.model_data, self.model_data_err, dropped = read_training_pickle(
            training_set_path, logTrafo, max_rel_uncertainty_data)
        self.design_points_org_ = self.design_points.copy()
        log.info("Training dataset size: %d, discarded points: %d", len(self.model_data), dropped)
        self.pardict = parse_model_parameter_file(parameter_file)
        bounds = np.array([[v[1], v[2]] for v in self.pardict.values()], dtype=np.float64)
        self.design_min, self.design_max = bounds[:, 0].copy(), bounds[:, 1].copy()
        self.nev, self.nobs = self.model_data.shape
        self.nparameters = self.design_points.shape[1]
        self.emu = None
        self._state = None
        self._device = None
        if self.parameterTrafoPCA_:
            self._fit_param_trafo()
            self.nparameters = self.PCA_new_design_points.shape[1]

    # ---- training: surmise, as in the reference -----------------------------------------------
    def trainEmulator(self, event_mask):
        try:
            from surmise.emulation import emulator
        except ImportError as exc:
            raise ImportError("training an EmulatorBAND needs the surmise package (surmise==0.2.1 in the "
                              "reference's requirements.txt); a trained emulator, or its fit information via "
                              "EmulatorBAND.from_fitinfo, can be evaluated without it") from exc
        mask = np.asarray(event_mask, dtype=bool)
        theta = (self.PCA_new_design_points if self.parameterTrafoPCA_ else self.design_points)[mask, :]
        log.info("Train GP emulators with %d training points ...", int(mask.sum()))
        x = np.arange(self.nobs).reshape(-1, 1)
        f = self.model_data[mask, :].T
        args = {"warnings": True}
        if self.method_ == "PCSK":
            args["simsd"] = self.model_data_err[mask, :].T
        method = "PCGPwImpute" if self.method_ == "PCGPwM" else self.method_   # as the reference maps it
        self.emu = emulator(x=x, theta=theta, f=f, method=method, args=args)
        self._state = self._device = None

    # ---- state / device ------------------------------------------------------------------------
    @classmethod
    def from_fitinfo(cls, info, design_min=None, design_max=None, exp_and_cov_diagonal=False):
        """A predict-only EmulatorBAND around surmise fit information (a dict shaped like `emu._info`)."""
        self = cls.__new__(cls)
        self.method_ = str(info.get("method", "PCGP"))
        self.logTrafo_ = bool(exp_and_cov_diagonal)
        self.parameterTrafoPCA_ = False
        self.exp_and_cov_diagonal_ = bool(exp_and_cov_diagonal)
        self.perform_no_PCA_ = False
        self.emu = info
        self.design_min, self.design_max = design_min, design_max
        self.nobs = int(np.asarray(info["offset"]).reshape(-1).shape[0])
        self.nev, self.nparameters = np.asarray(info["theta"]).shape
        self._state = self._device = None
        return self

    @property
    def state(self) -> EmulatorState:
        if self._state is None:
            if self.emu is None:
                raise RuntimeError("emulator is not trained")
            self._state = EmulatorState.from_trained(self)
        return self._state

    def _dev(self) -> DeviceEmulator:
        if self._device is None:
            self._device = DeviceEmulator(self.state)
        return self._device

    # ---- the hot path ------------------------------------------------------------------------
    def predict(self, X, return_cov=True, extra_std=0.0):
        """Model output at X [nsamples, nparameters]: mean [nsamples, nobs] and, with `return_cov`,
        cov [nsamples, nobs, nobs].  `extra_std` is accepted and ignored, as in the reference
        (src/emulator_BAND.py:386-478)."""
        return self._dev().predict(X, return_cov=return_cov, extra_std=0)

    def predict_diag(self, X, extra_std=0.0):
        return self._dev().predict_diag(X, extra_std=0)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_device"] = None
        return d
