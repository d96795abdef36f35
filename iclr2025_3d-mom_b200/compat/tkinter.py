W = "w"
