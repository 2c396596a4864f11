"""CPU: XML configuration semantics (KF.cpp:749-893) through kfpos_config_load_xml, checked
on the reference's shipped config strings (embedded verbatim: /root/reference is not available
at test time on the GPU box)."""
import pytest

from roskfpos_b200.batch import make_config

PX4 = '<config>\n  <px4flow armP0="1" armP1="0" useFixedSensorHeight="1" sensorHeight="5" sensorInitAngle ="-1.570796326794897" covarianceVelocity="0.04" covarianceGyroZ="0.02"/>\n</config>'
UWB = '<config>\n\t<!-- <uwb useFixedHeight="0" fixedHeight="1.049" tagId="6e5b"/> -->\n\t<uwb useFixedHeight="0" fixedHeight="1.049" tagId="0"/>\n</config>'
IMU = '<config>\n <imu useFixedCovarianceAcceleration="1" covarianceAcceleration="0.003"  useFixedCovarianceAngularVelocityZ="1" covarianceAngularVelocityZ="0.089"/> \n</config>'
MAG = '<config>\n<mag angleOffset="0" covarianceMag="0.0001"/>\n</config>'
POS = ('<config>\n <!--\n\tAlgorithm\n\ttype= 0: ML\n -->\n   <algorithm type="0" variant="1" numIgnoredRangings="2" '
       'bestMode="0" minZ="0.0" maxZ="3.0"  useInitPosition="1" initX="1.0" initY="1.0" initZ="1.048"/>\n</config>')


def test_shipped_configs(kflib):
    cfg = make_config(xml=[PX4, UWB, IMU, MAG, POS])
    assert (cfg.px4_arm_p0, cfg.px4_arm_p1, cfg.px4_sensor_height) == (1.0, 0.0, 5.0)
    assert cfg.px4_use_fixed_sensor_height == 1 and abs(cfg.px4_sensor_init_angle + 1.570796326794897) < 1e-15
    assert (cfg.px4_cov_velocity, cfg.px4_cov_gyro_z) == (0.04, 0.02)
    assert cfg.use_fixed_height == 0 and cfg.fixed_height == 1.049 and cfg.tag_id == 0
    assert cfg.imu_use_fixed_cov_acc == 1 and cfg.imu_cov_acc == 0.003
    assert cfg.imu_use_fixed_cov_gyro_z == 1 and cfg.imu_cov_gyro_z == 0.089
    assert cfg.mag_angle_offset == 0.0 and cfg.mag_cov == 0.0001
    assert cfg.variant == 1 and cfg.num_ignored_rangings == 2 and cfg.max_z == 3.0
    assert list(cfg.ml_start) == [1.0, 1.0, 1.048]


def test_defaults_and_unparsable_values(kflib):
    cfg = make_config(xml='<config><uwb fixedHeight="abc" tagId="6e5b"/><imu/></config>')
    assert cfg.fixed_height == 0.0 and cfg.tag_id == 0  # get<T>(path, 0) falls back to the default
    assert cfg.imu_use_fixed_cov_acc == 0 and cfg.imu_cov_acc == 0.0
    cfg = make_config(xml='<config><uwb useFixedHeight="2"/></config>')
    assert cfg.use_fixed_height == 0  # (use == 1) only, KF.cpp:796


def test_malformed_xml_is_reported(kflib):
    for bad in ("<config><mag angleOffset=0.1/></config>", "<nope/>", "<config><mag"):
        with pytest.raises(kflib.KfposError) as e:
            make_config(xml=bad)
        assert e.value.code == -6
