from .ode_model import OdeModel
from .abstract_model import ModelABC

__all__ = ['OdeModel', 'ModelABC']
