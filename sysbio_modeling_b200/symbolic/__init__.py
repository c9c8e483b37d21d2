from .sympy_tools import (parse_model_file, process_model_dict, symbolic_model_from_file,
                          symbolic_model_from_dict, trace_rhs_callable, emit_python, emit_cuda,
                          make_ode_model, SymbolicModel, CudaModelSource, count_flops)

__all__ = ['parse_model_file', 'process_model_dict', 'symbolic_model_from_file',
           'symbolic_model_from_dict', 'trace_rhs_callable', 'emit_python', 'emit_cuda',
           'make_ode_model', 'SymbolicModel', 'CudaModelSource', 'count_flops']
