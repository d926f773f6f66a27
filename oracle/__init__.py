"""ORACLE -- test infrastructure only (see oracle/README.md).  Never imported by the product
package; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it."""
