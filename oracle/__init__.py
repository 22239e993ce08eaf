"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the malstroem raster hot path (oracle/ms_oracle.c, loaded by oracle/port.py) and a
loader for the reference's own compiled Cython modules (oracle/ref.py -> oracle/_ref/*.so).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (malstroem_b200/) never does.
"""
