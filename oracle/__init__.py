"""CPU oracle for the BDPT hot path — test infrastructure only (see oracle/trace_ref.c)."""
