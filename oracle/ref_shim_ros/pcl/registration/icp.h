// unused by the member functions under test: empty stand-in (test infrastructure only)
