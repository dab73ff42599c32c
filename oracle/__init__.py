"""oracle -- CPU checkers for the parity tests.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Two checkers:

  port  oracle/libfm_oracle.so        plain-C restatement (fm_oracle.c)
  ref   oracle/_ref/libfm_ref.so      the UNMODIFIED reference headers behind a
                                      C shim (ref_harness.cpp), built in place from
                                      /root/reference by oracle/Makefile
"""
from .binding import Port, Ref, build, have_ref  # noqa: F401
