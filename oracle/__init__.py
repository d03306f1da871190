"""CPU oracle (test infrastructure only).  See oracle/rsb_oracle.c header: PARITY UNPINNED."""
