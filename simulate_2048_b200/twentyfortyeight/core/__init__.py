"""Game logic of the 2048 board: ``gameboard`` (moves, spawns, chance expansion) and
``gamemove`` (legal-action queries) — mirrors ``twentyfortyeight/core`` of the reference."""
