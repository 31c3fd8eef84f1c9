"""TEST INFRASTRUCTURE -- not product code.

CPU restatement (fp32 PyTorch / numpy) of the reference algorithm on the InternVLA-N1 hot path, used only as the
checker: by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  Nothing under
internnav_b200/ imports this package.
"""
