# placeholder, replaced below
