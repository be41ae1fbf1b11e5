"""Container-only stand-in for ``torchrl`` (TEST INFRASTRUCTURE). See README.md."""
