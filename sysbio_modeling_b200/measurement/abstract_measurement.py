"""Measurement base type (host data class, API kept from the reference).

Mirrors reference ``measurement/abstract_measurement.py:8-24``: a named vector of
values with per-point standard deviations; sigma == 0 is rejected because every
residual is divided by it (``squared_loss_function.py:40``).
"""
import numpy as np


class MeasurementABC(object):
    def __init__(self, variable_name, measurement_value, measurement_std=None):
        values = np.asarray(measurement_value, dtype=np.float64)
        if measurement_std is None:
            std = np.ones_like(values)
        else:
            std = np.asarray(measurement_std, dtype=np.float64)
        if np.any(std == 0):
            raise ValueError('Standard deviation of measurement cannot be 0')
        if len(values) != len(std):
            raise ValueError('Length of Standard Deviation Array Not Equal to Length of Measurements')
        self.variable_name = variable_name
        self.values = values
        self.std = std
