"""CPU oracle for the BioShell all-vs-all global-alignment hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this package.  The product
(bioshell_b200, libbioshell_align.so) never does.

* ``c_oracle``  -- ctypes binding of oracle/bioshell_oracle.c (literal C restatement)
* ``pyoracle``  -- independent pure-Python restatement used to cross-check it
"""
