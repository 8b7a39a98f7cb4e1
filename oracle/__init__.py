"""CPU oracle for the MultiAgentTracking step path -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package ``mate_b200`` never imports this.
"""
