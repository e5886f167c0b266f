"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms for the quantized Whisper decoder hot path.
Nothing under the product package may import this; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do (as the checker or the reported baseline).
"""
