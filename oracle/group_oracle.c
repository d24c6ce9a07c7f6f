/* group solvers: added below */
