"""Empty stand-in: the reference imports matplotlib at module top for plotting only."""
