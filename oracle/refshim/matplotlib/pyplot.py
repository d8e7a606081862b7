"""Empty stand-in for matplotlib.pyplot (plotting is never exercised)."""
