bash profiles/ab_generic.sh "- 3" "r64 3" "g8 3" "g8r64 3" "g4r64 3" "r64 3 HIVE_B200_STORE_CTAS=3" "g8r64 3 HIVE_B200_STORE_CTAS=3"
