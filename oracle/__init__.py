"""CPU oracle for the MVDeTr hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
Nothing under mvdetr_b200/ does (tests/test_layout.py enforces it).

  cpu_oracle   ctypes binding of the plain-C restatement (msda_ref.c, warp_ref.c) -> numpy in / numpy out
  torch_port   torch-CPU restatement of the reference's own CPU-capable Python path (multi-threaded; used as the
               timed CPU baseline and as a second, independent checker)
"""
