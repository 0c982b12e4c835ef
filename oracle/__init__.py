"""CPU oracle for the SLFP hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (cnns_slfp_quantization_b200) never does.
"""
