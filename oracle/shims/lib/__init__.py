"""TEST INFRASTRUCTURE: stand-in for the reference's absent ``lib`` package (oracle side)."""
