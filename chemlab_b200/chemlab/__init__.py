"""Python-3 re-authoring of chemlab's driver layer (L5-L3 of SURVEY section 1) on top of chemlab_b200.espressopp:
same command line (`start_simulation.py @params`), same `.top/.gro/.cfg` formats, same constructor sequence."""
