"""ORACLE - test infrastructure only.

CPU restatement of the reference's k-point path (numpy) plus the scripts that pin it against the real
reference.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from here; the product package pysktb_b200 must not.
"""
