"""oracle/ -- TEST INFRASTRUCTURE ONLY: CPU restatement of the reference algorithm (ric_oracle.c), the recipe that
compiles the real reference into oracle/_ref/, and ctypes access to both.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package."""
