"""Experiment data class (reference ``experiment/experiments.py:4-140``).

Host-side input type of the hot path.  Additive over the reference: an optional
``initial_conditions`` vector (the reference always integrates from zeros,
``ode_model.py:151-152`` / ``base_project.py:512``; BASELINE configs C2-C4 need
non-zero y(0)).
"""
import numpy as np


class Experiment(object):
    def __init__(self, name, measurements, fixed_parameters=None, experiment_settings=None,
                 initial_conditions=None):
        if not name[0].isalnum():
            raise ValueError("Experiment names must start with a letter or number")
        self.name = name
        self.fixed_parameters = fixed_parameters
        self.settings = dict(experiment_settings) if experiment_settings is not None else {}
        # Additive: None -> zeros (reference behaviour)
        self.initial_conditions = None if initial_conditions is None else \
            np.asarray(initial_conditions, dtype=np.float64)
        self._measurements = []
        if hasattr(measurements, '__iter__'):
            for m in measurements:
                self.add_measurement(m)
        else:
            self.add_measurement(measurements)
        self.param_global_vector_idx = None

    @property
    def measurements(self):
        return self._measurements

    def drop_timepoint_zero(self, variable=None):
        for m in self._measurements:
            if variable is None or m.variable_name == variable:
                m.drop_timepoint_zero()

    def get_unique_timepoints(self, include_zero=False):
        """Sorted union of all measurement times (reference :71-96)."""
        t = np.unique(np.concatenate([m.timepoints for m in self._measurements]))
        if not include_zero:
            t = t[t != 0]
        return t

    def get_variable_measurements(self, variable_name):
        for m in self._measurements:
            if m.variable_name == variable_name:
                return m
        raise KeyError('%s not in measurements' % variable_name)

    def add_measurement(self, measurement):
        for existing in self._measurements:
            if existing.variable_name == measurement.variable_name:
                raise KeyError('%s already has timeseries data associated with this experiment' % self.name)
        self._measurements.append(measurement)
        # measurements are kept sorted by variable name: this fixes the row order (reference :140)
        self._measurements.sort(key=lambda m: m.variable_name)
