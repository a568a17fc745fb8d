"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the reference's hot path.

Nothing in the product package (``acmil_b200/``) may import from here.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use these modules, and there only as the checker or
the reported CPU baseline, never as the thing shipped or measured as ours.
"""
