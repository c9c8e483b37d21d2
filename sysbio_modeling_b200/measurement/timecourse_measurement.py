"""Time-course measurement (reference ``measurement/timecourse_measurement.py:6-45``)."""
import numpy as np

from .abstract_measurement import MeasurementABC


class TimecourseMeasurement(MeasurementABC):
    """Values of one observable at a list of time points (+ optional sigma).

    ``get_nonzero_measurements`` drops the t == 0 points: the reference never fits
    them (``base_project.py:309``, ``utils.py:18``) because y(0) does not depend on
    the parameters.
    """

    def __init__(self, variable_name, measurement_value, measurement_time, measurement_std=None):
        super(TimecourseMeasurement, self).__init__(variable_name, measurement_value, measurement_std)
        timepoints = np.asarray(measurement_time, dtype=np.float64)
        if len(self.values) != len(timepoints):
            raise ValueError('Length of Standard Deviation Array Not Equal to Length of Timepoints')
        self.timepoints = timepoints

    def _nonzero_mask(self):
        return self.timepoints != 0

    def drop_timepoint_zero(self):
        keep = self._nonzero_mask()
        self.values = self.values[keep]
        self.std = self.std[keep]
        self.timepoints = self.timepoints[keep]

    def get_nonzero_measurements(self):
        keep = self._nonzero_mask()
        return self.values[keep], self.std[keep], self.timepoints[keep]
