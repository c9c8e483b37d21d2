from .timecourse_measurement import TimecourseMeasurement
from .abstract_measurement import MeasurementABC

__all__ = ['TimecourseMeasurement', 'MeasurementABC']
