"""CPU oracle for the CXL-SpecKV hot path.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py)."""
