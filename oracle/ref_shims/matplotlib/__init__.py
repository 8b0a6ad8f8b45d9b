"""Import shim — TEST INFRASTRUCTURE ONLY: lets oracle/make_golden.py import the reference's
my_environment/wrappers/wrappers.py (which imports pyplot at module level) without matplotlib."""
