"""State augmentation for temporally aggregated / mixed-frequency observations.

Host-side bookkeeping of ``DSGEStateSpace`` (gEconpy/model/statespace.py:556-723, 260-296, 334-388), kept in the
reference's vocabulary: *cumulator* states are deterministic lag copies of a "sum"/"mean"-aggregated variable, so that a
low-frequency observation ``y = w (x_t + x_{t-1} + ... + x_{t-s+1})`` is a linear function of the augmented state;
"first"/"last" aggregation needs no extra state, only missing values in the data (``prepare_mixed_frequency_data``).

    T_aug = [[T, 0], [F, kron(I, shift)]]      R_aug = [[R], [0]]      Z[i] = w e_var + w sum(e_slots)

Nothing here touches the device: the constant rows ``[F, kron(I, shift)]`` are written ONCE into the workspace the solver
kernel fills (``gecon_cr_args.t_ld / t_stride / r_stride``), so augmentation adds no pass over HBM per draw.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

CUMULATOR_AGGREGATIONS = ("sum", "mean")  # statespace.py:48
VALID_AGGREGATIONS = ("sum", "mean", "first", "last")


@dataclass
class StateAugmentation:
    """Layout of the augmented state vector.

    ``state_names``: names of the states the filter runs on, in filter order (the un-augmented part);
    ``observed_states``: data-column order; ``temporal_aggregation``: {observed state: "sum"|"mean"|"first"|"last"}.
    """

    state_names: list
    observed_states: list
    temporal_aggregation: dict = field(default_factory=dict)
    aggregation_period: int = 4
    # observation equations (statespace.py:390-556, 1040-1076): names of the observed series defined by an equation, and
    # the lag depth each model variable needs in the observation-lag block (already including the aggregation headroom)
    obs_equation_names: tuple = ()
    obs_lag_depths: dict = field(default_factory=dict)

    def __post_init__(self):
        ta = dict(self.temporal_aggregation or {})
        unknown = [v for v in ta if v not in self.observed_states]
        if unknown:
            raise ValueError(f"The following temporal_aggregation entries are not in observed_states: {', '.join(unknown)}")
        bad = {v: m for v, m in ta.items() if m not in VALID_AGGREGATIONS}
        if bad:
            raise ValueError(f"Unknown temporal aggregation method(s) {bad}; expected one of {VALID_AGGREGATIONS}")
        if any(m in CUMULATOR_AGGREGATIONS for m in ta.values()) and self.aggregation_period < 2:
            raise ValueError(f"aggregation_period must be >= 2 for sum/mean aggregation, got {self.aggregation_period}")
        missing = [v for v in self.observed_states if v not in self.state_names and v not in self.obs_equation_names]
        if missing:
            raise ValueError(f"observed states {missing} are not among the filter states")
        missing = [v for v in self.obs_lag_depths if v not in self.state_names]
        if missing:
            raise ValueError(f"observation equations reference {missing}, which are not among the filter states")
        self.temporal_aggregation = ta
        # fixed-order layout of the observation-lag block: each variable's slots run consecutively, in insertion order
        self._obs_lag_starts = {}
        offset = len(self.state_names) + self._n_cumulator_states
        for v, depth in self.obs_lag_depths.items():
            self._obs_lag_starts[v] = offset
            offset += int(depth)

    # ---- the reference's private properties, same names (statespace.py:556-584)
    @property
    def _cumulator_variables(self) -> list:
        # aggregation of an observation equation lives in the observation-lag block (statespace.py:562-571)
        return [v for v, m in self.temporal_aggregation.items() if m in CUMULATOR_AGGREGATIONS and v not in self.obs_equation_names]

    @property
    def _n_obs_lag_states(self) -> int:
        return int(sum(self.obs_lag_depths.values()))

    @property
    def _obs_lag_state_names(self) -> list:
        return [f"{v}_obs_lag{k}" for v, depth in self.obs_lag_depths.items() for k in range(1, depth + 1)]

    def _obs_lag_column(self, var_name: str, lag: int) -> int:
        """Augmented-state column of ``var_name`` lagged by ``-lag`` (lag < 0) (statespace.py:593-596)."""
        return self._obs_lag_starts[var_name] + (-lag - 1)

    @property
    def _n_cumulator_states(self) -> int:
        return len(self._cumulator_variables) * (self.aggregation_period - 1)

    @property
    def _cumulator_state_names(self) -> list:
        return [f"{v}_cumulator_lag{lag}" for v in self._cumulator_variables for lag in range(1, self.aggregation_period)]

    @property
    def k_orig_states(self) -> int:
        return len(self.state_names)

    @property
    def k_states(self) -> int:
        return self.k_orig_states + self._n_cumulator_states + self._n_obs_lag_states

    @property
    def augmented_state_names(self) -> list:
        return list(self.state_names) + self._cumulator_state_names + self._obs_lag_state_names

    # ---- constant blocks
    def transition_rows(self) -> np.ndarray:
        """The rows appended below T: cumulator block ``[F | kron(I, shift) | 0]`` (statespace.py:629-650), then the
        observation-lag block ``[F_lag | 0 | C_lag]`` (statespace.py:652-694); shape (k_states - k_orig, k_states)."""
        n_lags = self.aggregation_period - 1
        n_cum = self._n_cumulator_states
        rows = np.zeros((n_cum + self._n_obs_lag_states, self.k_states))
        k0 = self.k_orig_states
        for pos, name in enumerate(self._cumulator_variables):
            rows[pos * n_lags, self.state_names.index(name)] = 1.0  # F: slot 1 copies the variable
            for j in range(1, n_lags):  # shift companion: slot j+1 copies slot j
                rows[pos * n_lags + j, k0 + pos * n_lags + j - 1] = 1.0
        for v, depth in self.obs_lag_depths.items():
            start = self._obs_lag_starts[v]
            rows[start - k0, self.state_names.index(v)] = 1.0
            for j in range(1, depth):
                rows[start - k0 + j, start + j - 1] = 1.0
        return rows

    def augment_transition(self, T: np.ndarray) -> np.ndarray:
        """numpy twin of ``_augment_transition`` for (..., k, k) arrays."""
        n_cum = self._n_cumulator_states + self._n_obs_lag_states
        if n_cum == 0:
            return T
        k0 = self.k_orig_states
        out = np.zeros((*T.shape[:-2], k0 + n_cum, k0 + n_cum))
        out[..., :k0, :k0] = T
        out[..., k0:, :] = self.transition_rows()
        return out

    def augment_selection(self, R: np.ndarray) -> np.ndarray:
        """numpy twin of ``_augment_selection`` (statespace.py:696-723)."""
        n_cum = self._n_cumulator_states + self._n_obs_lag_states
        if n_cum == 0:
            return R
        return np.concatenate([R, np.zeros((*R.shape[:-2], n_cum, R.shape[-1]))], axis=-2)

    def design_matrix(self) -> np.ndarray:
        """``_make_design_matrix`` on the selector path (statespace.py:279-296)."""
        n_lags = self.aggregation_period - 1
        Z = np.zeros((len(self.observed_states), self.k_states))
        cum = self._cumulator_variables
        for i, name in enumerate(self.observed_states):
            if name in self.obs_equation_names:
                continue  # parameter-dependent row: filled per draw (design_cells)
            j = self.state_names.index(name)
            m = self.temporal_aggregation.get(name)
            if m in CUMULATOR_AGGREGATIONS:
                w = 1.0 / self.aggregation_period if m == "mean" else 1.0
                start = self.k_orig_states + cum.index(name) * n_lags
                Z[i, j] = w
                Z[i, start : start + n_lags] = w
            else:
                Z[i, j] = 1.0
        return Z

    def is_selector(self) -> bool:
        return self._n_cumulator_states == 0 and not self.obs_equation_names

    def design_cells(self, name: str, coeffs: dict) -> dict:
        """Cells of the design-matrix row of an observation equation: ``{column: [(weight, key), ...]}`` where ``coeffs``
        maps ``key = (variable, lag)`` to the linearisation coefficient.  With sum / mean aggregation over s periods the
        coefficient at lag k contributes at the effective lags k, k-1, ..., k-(s-1) (statespace.py:308-322)."""
        m = self.temporal_aggregation.get(name)
        n_periods = self.aggregation_period if m in CUMULATOR_AGGREGATIONS else 1
        weight = 1.0 / n_periods if m == "mean" else 1.0
        cells: dict = {}
        for (v, lag) in coeffs:
            for dd in range(n_periods):
                eff = lag - dd
                col = self.state_names.index(v) if eff == 0 else self._obs_lag_column(v, eff)
                cells.setdefault(col, []).append((weight, (v, lag)))
        return cells

    @staticmethod
    def required_obs_lag_depths(coeff_keys: dict, temporal_aggregation: dict, aggregation_period: int) -> dict:
        """``_obs_lag_depths`` (statespace.py:1040-1053): deepest effective lag per variable, ``coeff_keys`` =
        {observed series: iterable of (variable, lag)}."""
        depths: dict = {}
        for obs_name, keys in coeff_keys.items():
            broadcast = aggregation_period - 1 if (temporal_aggregation or {}).get(obs_name) in CUMULATOR_AGGREGATIONS else 0
            for v, lag in keys:
                need = -lag + broadcast
                if need > 0:
                    depths[v] = max(depths.get(v, 0), need)
        return depths

    def intercept_scale(self) -> np.ndarray:
        """Per observed state: aggregation_period for "sum", 1 otherwise (statespace.py:382-386)."""
        return np.array(
            [float(self.aggregation_period) if self.temporal_aggregation.get(v) == "sum" else 1.0 for v in self.observed_states]
        )


def prepare_mixed_frequency_data(low_freq_data, high_freq: str, aggregation_period: int = 4, observation_position: str = "last"):
    """Low-frequency observations on the model's (higher-frequency) calendar: each observation sits at the first or last period of
    its window of ``aggregation_period`` model periods, every other period is NaN (the filter masks NaN rows), and the frame
    ends at the last observation.  Signature and placement rule of gEconpy/model/statespace.py:1432-1509; the placement is
    computed for all observations at once (one sorted search of the observation dates in the model calendar)."""
    import pandas as pd

    if observation_position not in ("first", "last"):
        raise ValueError("observation_position must be 'first' or 'last'")
    offset = 0 if observation_position == "first" else int(aggregation_period) - 1
    n_hf = len(low_freq_data) * int(aggregation_period)
    calendar = pd.date_range(start=low_freq_data.index.min(), periods=n_hf, freq=high_freq)
    window_start = calendar.searchsorted(low_freq_data.index, side="left")  # first model period on or after each observation date
    target = window_start + offset
    placed = target < n_hf                                                  # windows cut short by the end of the calendar are dropped
    grid = np.full((n_hf, low_freq_data.shape[1]), np.nan)
    values = low_freq_data.to_numpy(dtype=np.float64)
    for t, row in zip(target[placed], values[placed]):  # in observation order: a later observation wins a shared slot
        grid[t] = row
    seen = np.flatnonzero(~np.isnan(grid).all(axis=1))
    if seen.size:
        grid, calendar = grid[: seen[-1] + 1], calendar[: seen[-1] + 1]
    out = pd.DataFrame(grid, index=calendar, columns=list(low_freq_data.columns))
    out.index.freq = out.index.inferred_freq
    return out
