"""
Drop-in for the reference's `src/emulator.py` Emulator with `predict` on the B200.

What stays as in the reference (it is offline, one-off work that produces the state the hot path
consumes): reading the training pickle, StandardScaler -> PCA(whiten) and
the scikit-learn GaussianProcessRegressor fits (src/emulator.py:257-363, 378-415).

What is replaced: `predict` (src/emulator.py:465-605).  The trained quantities are extracted into
an `EmulatorState`, uploaded once, and every call runs kernels (a) pc_predict and (b)
backtransform through the C ABI (include/gpbt.h).  There is no CPU prediction path.
"""
from __future__ import annotations

import logging
import pickle

import numpy as np

from . import parse_model_parameter_file
from .device import DeviceEmulator
from .state import GPR_ALPHA, EmulatorState

log = logging.getLogger(__name__)


def read_training_pickle(path, log_trafo=False, max_rel_uncertainty=0.1):
    """Training set in the reference's format (src/emulator.py:378-415):
    {event_id: {"parameter": [p], "obs": [2, m]}}; row 0 of obs is the value, row 1 its statistical
    error.  Events are taken in ascending integer id; an event whose largest relative error exceeds
    `max_rel_uncertainty` is dropped.  With `log_trafo` the observables become log|y| and the errors
    relative errors.  Returns (design [nev, p], data [nev, m], err [nev, m], n_dropped)."""
    with open(path, "rb") as fh:
        events = pickle.load(fh)
    design, data, err, dropped = [], [], [], 0
    for key in sorted(events, key=int):
        val, sig = np.asarray(events[key]["obs"], dtype=np.float64)
        if np.abs(sig / (val + 1e-16)).max() > max_rel_uncertainty:
            dropped += 1
            continue
        design.append(np.asarray(events[key]["parameter"], dtype=np.float64))
        if log_trafo:
            data.append(np.log(np.abs(val) + 1e-30))
            err.append(np.abs(sig / (val + 1e-30)))
        else:
            data.append(val)
            err.append(sig)
    return (np.array(design), np.array(data), np.nan_to_num(np.abs(np.array(err))), dropped)


def curve_on_grid(kind, params, grid):
    """The three parametrised curves of the reference evaluated on a grid, vectorised over rows of
    `params` (src/emulator.py:100-124): kind 0 zeta/s(T; zeta_max, T_zeta0, sigma_plus, sigma_minus)
    at mu_B = 0, kind 1 eta/s(mu_B; eta_0, eta_2, eta_4), kind 2 y_loss(y_init; yloss_2, _4, _6)."""
    a = [np.asarray(params, dtype=np.float64)[:, i][:, None] for i in range(np.shape(params)[1])]
    x = np.asarray(grid, dtype=np.float64)[None, :]
    if kind == 0:
        sig = np.where(x < a[1], a[3], a[2])
        return a[0] * np.exp(-(x - a[1]) ** 2. / (2. * sig ** 2.))
    if kind == 1:
        return np.where((0. < x) & (x <= 0.2), a[0] + (a[1] - a[0]) * (x / 0.2),
                        np.where((0.2 < x) & (x < 0.4), a[1] + (a[2] - a[1]) * ((x - 0.2) / 0.2), a[2] + 0. * x))
    if kind == 2:
        return np.where((0. < x) & (x <= 2.), a[0] * (x / 2.),
                        np.where((2. < x) & (x < 4.), a[0] + (a[1] - a[0]) * ((x - 2.) / 2.),
                                 a[1] + (a[2] - a[1]) * ((x - 4.) / 2.)))
    raise ValueError(kind)


class Emulator:
    """PCA + independent-GP emulator; same constructor and public methods as the reference class.

    Attributes kept for callers / notebooks: design_min, design_max, npc, nev, nobs, model_data,
    model_data_err, design_points, scaler, pca, gps, _trans_matrix, _var_trans, _cov_trunc."""

    def __init__(self, training_set_path=".", parameter_file="ABCD.txt", npc=10, nrestarts=0,
                 logTrafo=False, parameterTrafoPCA=False, max_rel_uncertainty_data=0.1,
                 exp_and_cov_diagonal=False, perform_no_PCA=False):
        from sklearn.decomposition import PCA
        from sklearn.preprocessing import StandardScaler
        if exp_and_cov_diagonal and not logTrafo:
            raise ValueError("exp_and_cov_diagonal can only be set to True if logTrafo is True.")
        self.logTrafo_ = logTrafo
        self.parameterTrafoPCA_ = bool(parameterTrafoPCA)
        self.max_rel_uncertainty_data_ = max_rel_uncertainty_data
        self.exp_and_cov_diagonal_ = exp_and_cov_diagonal
        self.perform_no_PCA_ = perform_no_PCA
        self.npc = npc
        self.nrestarts = nrestarts

        self.design_points, self.model_data, self.model_data_err, dropped = read_training_pickle(
            training_set_path, logTrafo, max_rel_uncertainty_data)
        self.design_points_org_ = self.design_points.copy()
        log.info("Training dataset size: %d, discarded points: %d", len(self.model_data), dropped)
        self.nev, self.nobs = self.model_data.shape

        self.pardict = parse_model_parameter_file(parameter_file)
        bounds = np.array([[v[1], v[2]] for v in self.pardict.values()], dtype=np.float64)
        self.design_min, self.design_max = bounds[:, 0].copy(), bounds[:, 1].copy()

        self.scaler = StandardScaler(copy=False)
        self.pca = PCA(copy=False, whiten=True, svd_solver="full")
        self.gps = []
        self._state = None
        self._device = None
        if self.parameterTrafoPCA_:
            self._fit_param_trafo()

    # ---- parameter-function PCA (offline; the per-walker transform itself runs on the GPU) -----
    def _fit_param_trafo(self, target_variance=0.99):
        """Replace the zeta/s, eta/s and y_loss parameter groups of the design by the principal
        components (99 % of the variance) of the curves they parametrise, in the reference's order
        bulk -> shear -> y_loss, and move the design box along (src/emulator.py:79-99, 129-241)."""
        from sklearn.decomposition import PCA
        from sklearn.preprocessing import StandardScaler
        from .state import PARAM_TRAFO_GROUPS
        if self.design_points.shape[1] < 19:
            raise ValueError("parameterTrafoPCA needs the 19+ parameter layout of the reference "
                             "(columns 2-4, 12-14, 15-18 are the transformed groups)")
        self.targetVariance = target_variance
        self.indices_zeta_s_parameters = [15, 16, 17, 18]
        self.indices_eta_s_parameters = [12, 13, 14]
        self.indices_yloss_parameters = [2, 3, 4]
        new_design = self.design_points
        for kind, tag, idx_attr, grid in PARAM_TRAFO_GROUPS:
            idx = getattr(self, idx_attr)
            curves = curve_on_grid(kind, self.design_points[:, idx], np.linspace(*grid))
            scaler = StandardScaler()
            pca = PCA(n_components=target_variance)
            pcs = pca.fit(scaler.fit_transform(curves)).transform(scaler.transform(curves))
            setattr(self, "paramTrafoScaler_" + tag, scaler)
            setattr(self, "paramTrafoPCA_" + tag, pca)
            log.info("%s parameter PCA uses %d PCs", tag, pca.n_components_)
            # indices refer to the ORIGINAL columns: groups are removed back to front (15-18, 12-14,
            # 2-4), and the components are appended behind everything else
            new_design = np.concatenate((np.delete(new_design, idx, axis=1), pcs), axis=1)
            self.design_min = np.concatenate((np.delete(self.design_min, idx), pcs.min(axis=0)))
            self.design_max = np.concatenate((np.delete(self.design_max, idx), pcs.max(axis=0)))
        self.PCA_new_design_points = new_design

    # ---- training (offline, scikit-learn, as the reference) -----------------------------------
    def trainEmulatorAutoMask(self):
        self.trainEmulator([True] * self.nev)

    def _make_kernel(self, kernel_type):
        from sklearn.gaussian_process import kernels
        span = self.design_max - self.design_min
        if kernel_type == "RBF":
            base = kernels.RBF(length_scale=span, length_scale_bounds=np.outer(span, (1e-1, 1e2)))
        elif kernel_type == "Matern":
            base = kernels.Matern(length_scale=span, length_scale_bounds=np.outer(span, (1e-3, 1e5)), nu=1.5)
        else:
            raise ValueError("Unknown kernel type: {}".format(kernel_type))
        noise = kernels.WhiteKernel(noise_level=.05, noise_level_bounds=(1e-2, 1e2))
        return 1. * base + noise

    def trainEmulator(self, eventMask, kernel_type="RBF"):
        from sklearn.gaussian_process import GaussianProcessRegressor
        mask = np.asarray(eventMask, dtype=bool)
        Y = self.scaler.fit_transform(self.model_data[mask, :])
        if self.perform_no_PCA_:
            Z = Y
        else:
            Z = self.pca.fit_transform(Y)[:, :self.npc]
            log.info("%d PCs explain %.5f of variance", self.npc,
                     self.pca.explained_variance_ratio_[:self.npc].sum())
        theta = (self.PCA_new_design_points if self.parameterTrafoPCA_ else self.design_points)[mask, :]
        kernel = self._make_kernel(kernel_type)
        self.gps = [GaussianProcessRegressor(kernel=kernel, alpha=GPR_ALPHA,
                                             n_restarts_optimizer=self.nrestarts,
                                             copy_X_train=False).fit(theta, z) for z in Z.T]
        self._kernel_type = kernel_type
        self._finish_training()

    def _finish_training(self):
        """Transform matrices of the PCA mode (src/emulator.py:330-363)."""
        if not self.perform_no_PCA_:
            self._trans_matrix = (self.pca.components_
                                  * np.sqrt(self.pca.explained_variance_[:, np.newaxis])
                                  * self.scaler.scale_)
            kept, dropped = self._trans_matrix[:self.npc], self._trans_matrix[self.npc:]
            self._cov_trunc = dropped.T @ dropped
            self._cov_trunc.flat[::self.nobs + 1] += 1e-4 * self.scaler.var_
        self._state = None
        self._device = None

    @property
    def _var_trans(self):
        # the reference precomputes this [npc, nobs^2] array (src/emulator.py:353-355); kernel (b)
        # never needs it, so it is only materialised if a caller asks
        kept = self._trans_matrix[:self.npc]
        return np.einsum("ki,kj->kij", kept, kept).reshape(self.npc, self.nobs ** 2)

    # ---- state / device --------------------------------------------------------------------
    @property
    def state(self) -> EmulatorState:
        if self._state is None:
            if not self.gps:
                raise RuntimeError("emulator is not trained")
            self._state = EmulatorState.from_trained(self, keep_L=True)
        return self._state

    @classmethod
    def from_state(cls, state: EmulatorState, design_min=None, design_max=None):
        """An Emulator that only predicts (no training data), around an existing state."""
        self = cls.__new__(cls)
        self.logTrafo_ = state.exp_diag
        self.parameterTrafoPCA_ = state.trafo is not None
        self.exp_and_cov_diagonal_ = state.exp_diag
        self.perform_no_PCA_ = state.no_pca
        self.npc, self.nobs, self.nev = state.q, state.m, state.n
        self.design_min, self.design_max = design_min, design_max
        self.gps = []
        self._state, self._device = state, None
        return self

    def _dev(self) -> DeviceEmulator:
        if self._device is None:
            self._device = DeviceEmulator(self.state)
        return self._device

    # ---- the hot path ------------------------------------------------------------------------
    def predict(self, X, return_cov=True, extra_std=0):
        """Model output at X [nsamples, ndim]: mean [nsamples, nobs] and, with `return_cov`, the
        observable-space covariance [nsamples, nobs, nobs] per sample.  `extra_std` (scalar or one
        value per sample) is added in quadrature to every GP's predictive standard deviation."""
        return self._dev().predict(X, return_cov=return_cov, extra_std=extra_std)

    def predict_diag(self, X, extra_std=0):
        """(mean, var): `var` is the diagonal of the covariance `predict` returns, without forming
        the nsamples x nobs x nobs array -- for posterior-predictive and sensitivity sweeps."""
        return self._dev().predict_diag(X, extra_std=extra_std)

    # ---- small helpers the notebooks call (src/emulator.py:100-124, 243-248, 366-375, 608-633) ------
    def parametrization_zeta_over_s_vs_T(self, zeta_max, T_zeta0, sigma_plus, sigma_minus, T, mu_B):
        """zeta/s(T): Gaussian bump of height zeta_max around T_zeta0 - 0.15 mu_B^2, width sigma_minus
        below T_zeta0 and sigma_plus above (scalar or array T)."""
        centre = T_zeta0 - 0.15 * mu_B ** 2.
        width = np.where(np.asarray(T) < T_zeta0, sigma_minus, sigma_plus)
        out = zeta_max * np.exp(-(T - centre) ** 2. / (2. * width ** 2.))
        return float(out) if np.ndim(out) == 0 else out

    def parametrization_eta_over_s_vs_mu_B(self, eta_0, eta_2, eta_4, mu_B):
        out = curve_on_grid(1, [[eta_0, eta_2, eta_4]], np.atleast_1d(mu_B))[0]
        return float(out[0]) if np.ndim(mu_B) == 0 else out

    def parametrization_y_loss_vs_y_init(self, yloss_2, yloss_4, yloss_6, y_init):
        out = curve_on_grid(2, [[yloss_2, yloss_4, yloss_6]], np.atleast_1d(y_init))[0]
        return float(out[0]) if np.ndim(y_init) == 0 else out

    def outputPCAvsParam(self):
        """design points and the first npc principal-component scores of the training data [npc, nev]"""
        # (a copy: scaler and pca work in place, copy=False as in the reference)
        Z = self.pca.fit_transform(self.scaler.fit_transform(self.model_data.copy()))[:, :self.npc]
        return self.design_points, Z.T

    def _inverse_transform(self, Z):
        """principal components [..., k <= nobs] -> observables [..., nobs]"""
        return np.dot(Z, self._trans_matrix[:np.shape(Z)[-1]]) + self.scaler.mean_

    def sample_y(self, X, n_samples=1, random_state=None):
        """Draws of the model output at X, [n_X, n_samples, nobs]: each emulated PC is sampled from its
        GP (scikit-learn, on the host, as in the reference), the truncated components are standard
        normal.  Not available without the PCA (the reference has no such path either); X is in the
        GPs' own input space (the reference does not apply the parameterTrafoPCA transform here)."""
        if self.perform_no_PCA_:
            log.warning("Sampling from raw data is not implemented.")
            return None
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        draws = [gp.sample_y(X, n_samples=n_samples, random_state=random_state)[:, :, np.newaxis] for gp in self.gps]
        rest = np.random.standard_normal((X.shape[0], n_samples, self.pca.n_components_ - self.npc))
        return self._inverse_transform(np.concatenate(draws + [rest], axis=2))

    # ---- validation helpers (callers of predict; src/emulator.py:418-421, 636-726) ----------------
    def getAvgTrainingDataRelError(self):
        """mean over the design of (statistical error / value) per observable"""
        return np.mean(np.nan_to_num(self.model_data_err / self.model_data), axis=0)

    def _hold_out(self, nTestPoints, test_on_training_points):
        """Train without the last nTestPoints design points, predict either those (validation) or the
        points trained on (closure); returns (prediction, its std, data, data error), each
        [points, nobs], back on the linear scale when the emulator works with log observables.
        The std comes from predict_diag: same numbers as sqrt(diag(cov)) without the covariances."""
        log.info("Validating GP emulator ...")
        train = np.ones(self.nev, dtype=bool)
        train[self.nev - nTestPoints:] = False
        self.trainEmulator(list(train))
        rows = train if test_on_training_points else ~train
        pred, var = self.predict_diag(self.design_points_org_[rows, :])
        std = np.sqrt(var)
        if self.logTrafo_ and not self.exp_and_cov_diagonal_:
            pred, std = np.exp(pred), std * np.exp(pred)
        data, err = self.model_data[rows, :], self.model_data_err[rows, :]
        if self.logTrafo_:
            data, err = np.exp(data), err * np.exp(data)
        shape = (-1, self.nobs)
        return pred.reshape(shape), std.reshape(shape), data.reshape(shape), err.reshape(shape)

    def testEmulatorErrors(self, nTestPoints=1):
        """Train on all but the last nTestPoints design points and predict those."""
        return self._hold_out(nTestPoints, test_on_training_points=False)

    def testEmulatorErrorsWithTrainingPoints(self, nTestPoints=1):
        """Same training, but the predictions are made at the points trained on."""
        return self._hold_out(nTestPoints, test_on_training_points=True)

    # ---- pickling: device handles never travel -------------------------------------------------
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_device"] = None
        return d
