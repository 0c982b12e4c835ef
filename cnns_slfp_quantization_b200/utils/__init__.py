"""Mirror of the reference's `utils` package (utils/sfp_quant.py, conv2d_func.py,
activation_func.py, optimizer.py): same module names, callables and signatures."""
