"""ORACLE: CPU restatement of the reference's G1/Fr hot path. Test infrastructure only.

Nothing under oracle/ may be imported by the product package (curdleproofs_pie_b200);
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
"""
