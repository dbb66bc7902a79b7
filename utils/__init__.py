"""`utils` package of the drop-in surface (utils/{buffer,model,recorder,runner,terrain,utils}.py of the reference)."""
