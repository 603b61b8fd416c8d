"""Stand-in for `matplotlib` (absent offline): imported by the reference's system / eval modules for plots only."""
