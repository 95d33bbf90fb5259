"""TEST INFRASTRUCTURE ONLY -- stand-in for the two fvcore symbols the reference uses."""
