"""``LinearElliptic::Estimators::SWIPDG`` / ``BlockSWIPDG`` string dispatchers
(estimators/swipdg.hh:824-985, estimators/block-swipdg.hh:1076-1265): static ``available()``, ``available_local()``,
``estimate(space, vector, problem, type[, parameters])`` and ``estimate_local(...)``.  The discretization object plays the
role of (space, problem)."""

ESV2007_TYPES = ["eta_NC_ESV2007", "eta_R_ESV2007", "eta_R_ESV2007_*", "eta_DF_ESV2007", "eta_ESV2007",
                 "eta_ESV2007_alt"]
OS2014_TYPES = ["eta_NC_OS2014", "eta_R_OS2014", "eta_R_OS2014_*", "eta_DF_OS2014", "eta_DF_OS2014_*", "eta_OS2014",
                "eta_OS2014_*"]


class SWIPDG:
    @staticmethod
    def available(discretization):
        return [t for t in discretization.available_estimators() if t in ESV2007_TYPES]

    @staticmethod
    def available_local(discretization):
        return [t for t in SWIPDG.available(discretization) if t in ("eta_ESV2007", "eta_ESV2007_alt")]

    @staticmethod
    def estimate(discretization, vector, type, parameters=None):
        if type not in ESV2007_TYPES:
            from .discretizations import you_are_using_this_wrong
            raise you_are_using_this_wrong(2, "Requested type '%s' is not one of available()!" % type)
        return discretization.estimate(vector, type, parameters)

    @staticmethod
    def estimate_local(discretization, vector, type, parameters=None):
        if type not in ("eta_ESV2007", "eta_ESV2007_alt"):
            from .discretizations import you_are_using_this_wrong
            raise you_are_using_this_wrong(2, "Requested type '%s' is not one of available_local()!" % type)
        return discretization.estimate_local(vector, type, parameters)


class BlockSWIPDG:
    @staticmethod
    def available(discretization):
        return [t for t in discretization.available_estimators() if t in OS2014_TYPES]

    @staticmethod
    def available_local(discretization):
        return [t for t in BlockSWIPDG.available(discretization) if t in ("eta_OS2014", "eta_OS2014_*")]

    @staticmethod
    def estimate(discretization, vector, type, parameters=None):
        if type not in OS2014_TYPES:
            from .discretizations import you_are_using_this_wrong
            raise you_are_using_this_wrong(2, "Requested type '%s' is not one of available()!" % type)
        return discretization.estimate(vector, type, parameters)

    @staticmethod
    def estimate_local(discretization, vector, type, parameters=None):
        if type not in ("eta_OS2014", "eta_OS2014_*"):
            from .discretizations import you_are_using_this_wrong
            raise you_are_using_this_wrong(2, "Requested type '%s' is not one of available_local()!" % type)
        return discretization.estimate_local(vector, type, parameters)
