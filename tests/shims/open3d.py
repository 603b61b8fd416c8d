"""Stand-in for `open3d` (absent offline): imported by /root/reference/src/system.py, never called in the loop."""
