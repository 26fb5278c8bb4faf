"""ase.data stand-in: atomic_numbers / chemical_symbols. Test infrastructure only."""
chemical_symbols = ["X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si",
                    "P", "S", "Cl", "Ar", "K", "Ca"]
atomic_numbers = {s: z for z, s in enumerate(chemical_symbols)}
