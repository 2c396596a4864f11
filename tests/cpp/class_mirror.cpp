// Drives the header-only mirror classes (include/kfpos/PositionEstimationAlgorithm.hpp) the way
// PosGenerator drives the reference classes, with a scripted input read from stdin, and prints
// the resulting poses.  Built and checked against the oracle by tests/test_gpu_cpp_classes.py.
#include <cstdio>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

#include <kfpos/PositionEstimationAlgorithm.hpp>

using namespace kfpos;

int main() {
    std::unique_ptr<PositionEstimationAlgorithm> algo;
    std::vector<Beacon> beacons;
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        std::string cmd;
        in >> cmd;
        if (cmd == "anchors") {
            int n;
            in >> n;
            beacons.clear();
            for (int i = 0; i < n; ++i) {
                Beacon b;
                b.id = i; b.index = i;
                in >> b.position.x >> b.position.y >> b.position.z;
                beacons.push_back(b);
            }
        } else if (cmd == "t6") {
            double a, thr, x, y, z; int loo;
            in >> a >> loo >> thr >> x >> y >> z;
            algo.reset(new KalmanFilterTOA(a, loo != 0, thr, Vector3(x, y, z)));
        } else if (cmd == "t9") {
            double a, j, x, y, z;
            in >> a >> j >> x >> y >> z;
            algo.reset(new KalmanFilterTOAIMU(a, j, Vector3(x, y, z)));
        } else if (cmd == "k8") {
            double a, ang, j, x, y;
            in >> a >> ang >> j >> x >> y;
            const std::string px4 = "<config><px4flow armP0=\"1\" armP1=\"0\" sensorHeight=\"5\" covarianceVelocity=\"0.04\" covarianceGyroZ=\"0.02\"/></config>";
            const std::string tag = "<config><uwb useFixedHeight=\"0\" fixedHeight=\"1.049\" tagId=\"0\"/></config>";
            const std::string imu = "<config><imu useFixedCovarianceAcceleration=\"1\" covarianceAcceleration=\"0.003\" useFixedCovarianceAngularVelocityZ=\"1\" covarianceAngularVelocityZ=\"0.089\"/></config>";
            const std::string mag = "<config><mag angleOffset=\"0\" covarianceMag=\"0.0001\"/></config>";
            const std::string pos = "<config><algorithm type=\"1\" variant=\"0\"/></config>";
            algo.reset(new KalmanFilter(a, ang, j, pos, px4, tag, imu, mag, Vector3(x, y, 0)));
            if (!algo->init()) { std::printf("init failed\n"); return 2; }
        } else if (cmd == "t9_nofix") {
            double a, j;
            in >> a >> j;
            algo.reset(new KalmanFilterTOAIMU(a, j));
        } else if (cmd == "k8_nofix") { // the constructor without initialPosition; fh = useFixedHeight
            double a, ang, j; int fh;
            in >> a >> ang >> j >> fh;
            const std::string px4 = "<config><px4flow armP0=\"1\" armP1=\"0\" sensorHeight=\"5\" covarianceVelocity=\"0.04\" covarianceGyroZ=\"0.02\"/></config>";
            const std::string tag = std::string("<config><uwb useFixedHeight=\"") + (fh ? "1" : "0") + "\" fixedHeight=\"1.049\" tagId=\"0\"/></config>";
            const std::string imu = "<config><imu useFixedCovarianceAcceleration=\"1\" covarianceAcceleration=\"0.003\" useFixedCovarianceAngularVelocityZ=\"1\" covarianceAngularVelocityZ=\"0.089\"/></config>";
            const std::string mag = "<config><mag angleOffset=\"0\" covarianceMag=\"0.0001\"/></config>";
            const std::string pos = "<config><algorithm type=\"1\" variant=\"0\"/></config>";
            algo.reset(new KalmanFilter(a, ang, j, pos, px4, tag, imu, mag));
            if (!algo->init()) { std::printf("init failed\n"); return 2; }
        } else if (cmd == "ml") {
            int use2d, variant, nign; double x, y, z;
            in >> use2d >> variant >> nign >> x >> y >> z;
            algo.reset(new MLLocation(use2d != 0, variant, nign, Vector3(x, y, z)));
        } else if (cmd == "toa") {
            double dt, err;
            in >> dt >> err;
            std::vector<double> r, e;
            std::vector<Beacon> sel;
            for (size_t i = 0; i < beacons.size(); ++i) {
                double v;
                in >> v;
                if (v > 0) { // PosGenerator forwards only valid slots (PG.cpp:481-489)
                    r.push_back(v); e.push_back(err); sel.push_back(beacons[i]);
                }
            }
            algo->setNextDt(dt);
            algo->newTOAMeasurement(r, sel, e, 0.0);
        } else if (cmd == "imu") {
            double dt; VectorDim3 w, a; double cw[9], ca[9];
            in >> dt >> w.x >> w.y >> w.z >> a.x >> a.y >> a.z;
            for (int i = 0; i < 9; ++i) in >> cw[i];
            for (int i = 0; i < 9; ++i) in >> ca[i];
            algo->setNextDt(dt);
            algo->newIMUMeasurement(w, cw, a, ca);
        } else if (cmd == "px4") {
            double dt, ix, iy, irz, it; int q;
            in >> dt >> ix >> iy >> irz >> it >> q;
            algo->setNextDt(dt);
            algo->newPX4FlowMeasurement(ix, iy, irz, it, q);
        } else if (cmd == "compass") {
            double dt, c;
            in >> dt >> c;
            algo->setNextDt(dt);
            algo->newCompassMeasurement(c);
        } else if (cmd == "mag") {
            double dt; VectorDim3 m; double c[9] = {0};
            in >> dt >> m.x >> m.y >> m.z;
            algo->setNextDt(dt);
            algo->newMAGMeasurement(m, c);
        } else if (cmd == "pose") {
            Vector3 p;
            const bool ok = algo->getPose(p);
            std::printf("pose %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d", ok ? 1 : 0, p.x, p.y, p.z,
                        p.rotZ, p.rotW, p.linearSpeedX, p.linearSpeedY, p.angularSpeedZ,
                        p.covarianceMatrix.n_rows ? p.covarianceMatrix(0, 0) : 0.0, algo->status());
            std::printf("\n");
        }
    }
    return 0;
}
