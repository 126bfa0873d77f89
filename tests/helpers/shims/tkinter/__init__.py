E = "e"                       # dataset/dataset_deform4d_flow.py:5 `from tkinter import E` (unused)
