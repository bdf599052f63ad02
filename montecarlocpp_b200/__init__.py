"""B200-native hot path of nickdou/montecarlocpp (FieldProblem::solve) behind a C ABI.

The product is libmcb.so (montecarlocpp_b200/csrc, include/mcb.h) plus the C++ host mirror
of the reference's Problem/Domain/Material API (montecarlocpp_b200/host).  This Python
package is plumbing for tests and bench.py: ctypes bindings and synthetic material files.
"""
__all__ = ["abi", "materials"]
